// align_reader.cpp -- SAM text / BAM decode straight into structure-of-arrays batches, and BAM re-emission.
//
// Replaces the reference's use of htslib (sam_open / sam_hdr_read / sam_read1 / sam_write1) around the hot path:
// only the fields the path consumes are materialised (tid, pos, flag, l_qseq, CIGAR words, NM, XS:A, qname), in the
// layout of lrb_batch.  Field semantics follow the BAM spec and htslib's parser for this submodule version
// (htslib/sam.c:406-440 bam_read1, :844-1041 sam_parse1, :1280-1360 aux accessors) -- e.g. SAM-text normalisations
// "tid<0 => FUNMAP", "CIGAR '*' => FUNMAP", integer aux tags stored in the smallest fitting type.
#include <cctype>
#include <cerrno>
#include <cstdlib>
#include <cstring>
#include <algorithm>
#include <atomic>
#include <thread>
#include <fcntl.h>
#include <sys/stat.h>
#include <unistd.h>
#include <zlib.h>
#include "lrb_host.h"

namespace lrb {

// ---- host threads (the reference decodes on one thread; htslib's own answer is hts_set_threads, hts.c:969)
// LRB_THREADS=n overrides; 1 gives the purely sequential code paths below.
int host_threads()
{
    static int n = 0;
    if (!n) {
        const char *e = getenv("LRB_THREADS");
        n = e ? atoi(e) : (int)std::thread::hardware_concurrency();
        if (n < 1) n = 1;
        if (n > 64) n = 64;
    }
    return n;
}
// fn(i) for i in [0, n) on up to host_threads() threads (dynamic, one index at a time: the work items are coarse)
void parallel_for(size_t n, const std::function<void(size_t)> &fn)
{
    size_t nt = (size_t)host_threads(); if (nt > n) nt = n;
    if (nt <= 1) { for (size_t i = 0; i < n; ++i) fn(i); return; }
    std::atomic<size_t> next{0};
    auto work = [&] { for (size_t i; (i = next.fetch_add(1)) < n;) fn(i); };
    std::vector<std::thread> th;
    for (size_t t = 1; t < nt; ++t) th.emplace_back(work);
    work();
    for (auto &t : th) t.join();
}

uint64_t hash_name(const char *s, size_t n)
{
    // 64-bit multiply-xorshift hash over 8-byte words (seeded with the length); names differing anywhere differ in
    // hash with probability 1-2^-64, which is what lrb_batch.qname_hash asks for.
    uint64_t h = 0x9E3779B97F4A7C15ull ^ (n * 0xD6E8FEB86659FD93ull);
    size_t i = 0;
    for (; i + 8 <= n; i += 8) {
        uint64_t w; memcpy(&w, s + i, 8);
        h ^= w; h *= 0xFF51AFD7ED558CCDull; h ^= h >> 32;
    }
    uint64_t w = 0;
    if (i < n) memcpy(&w, s + i, n - i);
    h ^= w; h *= 0xC4CEB9FE1A85EC53ull; h ^= h >> 29;
    h *= 0xFF51AFD7ED558CCDull; h ^= h >> 32;
    return h;
}

lrb_batch Records::view() const
{
    lrb_batch b;
    b.n = (int64_t)n(); b.tid = tid.data(); b.pos = pos.data(); b.flag = flag.data(); b.l_qseq = l_qseq.data();
    b.nm = nm.data(); b.xs = xs.data(); b.qname_hash = qhash.data(); b.cigar_off = cigar_off.data(); b.cigar = cigar.data();
    return b;
}

static bool slurp(const std::string &path, Bytes &buf, std::string &err)
{
    if (path != "-") {                                                 // regular file: sized once, read by all threads (pread)
        int fd = open(path.c_str(), O_RDONLY);
        if (fd < 0) { err = "Cannot open \"" + path + "\""; return false; }
        struct stat st;
        if (fstat(fd, &st) == 0 && S_ISREG(st.st_mode) && st.st_size > 0) {
            const size_t n = (size_t)st.st_size, CH = 8u << 20, nch = (n + CH - 1) / CH;
            buf.reserve(n + 1); buf.resize(n);                          // one spare byte: parse_sam terminates the last line in place
            std::atomic<bool> bad{false};
            parallel_for(nch, [&](size_t k) {
                size_t o = k * CH, e = o + CH < n ? o + CH : n;
                while (o < e) { ssize_t g = pread(fd, buf.data() + o, e - o, (off_t)o); if (g <= 0) { bad = true; return; } o += (size_t)g; }
            });
            close(fd);
            if (bad) { err = "Cannot read \"" + path + "\""; return false; }
            return true;
        }
        close(fd);
    }
    FILE *fp = (path == "-") ? stdin : fopen(path.c_str(), "rb");
    if (!fp) { err = "Cannot open \"" + path + "\""; return false; }
    size_t cap = 1 << 20, n = 0;
    buf.resize(cap);
    for (;;) {
        size_t k = fread(buf.data() + n, 1, cap - n, fp);
        n += k;
        if (k == 0) break;
        if (n == cap) { cap *= 2; buf.resize(cap); }
    }
    buf.resize(n);
    if (fp != stdin) fclose(fp);
    return true;
}

bool read_file_bytes(const std::string &path, Bytes &buf) { std::string e; return path != "-" && slurp(path, buf, e); }

// concatenated gzip members (BGZF blocks are gzip members, bgzf.c) -> one buffer
static bool gunzip_all(const Bytes &in, Bytes &out, std::string &err)
{
    z_stream zs; memset(&zs, 0, sizeof zs);
    if (inflateInit2(&zs, 15 + 32) != Z_OK) { err = "zlib init failed"; return false; }
    out.resize(in.size() * 4 + (1 << 16));
    zs.next_in = (Bytef *)in.data(); zs.avail_in = (uInt)0;
    size_t ipos = 0, opos = 0;
    while (ipos < in.size()) {
        size_t chunk = in.size() - ipos; if (chunk > (1u << 30)) chunk = 1u << 30;
        zs.next_in = (Bytef *)in.data() + ipos; zs.avail_in = (uInt)chunk;
        for (;;) {
            if (out.size() - opos < (1 << 16)) out.resize(out.size() * 2);
            size_t room = out.size() - opos; if (room > (1u << 30)) room = 1u << 30;
            zs.next_out = out.data() + opos; zs.avail_out = (uInt)room;
            int rc = inflate(&zs, Z_NO_FLUSH);
            opos += room - zs.avail_out;
            if (rc == Z_STREAM_END) {
                size_t used = chunk - zs.avail_in;
                ipos += used; chunk -= used;
                inflateReset(&zs);
                if (chunk == 0) break;
                zs.next_in = (Bytef *)in.data() + ipos; zs.avail_in = (uInt)chunk;
                continue;
            }
            if (rc != Z_OK && rc != Z_BUF_ERROR) { inflateEnd(&zs); err = "corrupt gzip/BGZF stream"; return false; }
            if (zs.avail_in == 0 && zs.avail_out != 0) { ipos += chunk; chunk = 0; break; }   // truncated member
        }
    }
    inflateEnd(&zs);
    out.resize(opos);
    return true;
}

static inline uint32_t rd32(const uint8_t *p) { uint32_t v; memcpy(&v, p, 4); return v; }
static inline int32_t rdi32(const uint8_t *p) { int32_t v; memcpy(&v, p, 4); return v; }

// BGZF (bgzf.c block format): every block is a gzip member whose extra field carries 'BC' = block size - 1 and whose
// trailer carries the inflated size, so the blocks can be located without inflating and inflated independently, straight
// into their final place.  Returns false (out untouched) when the stream is not pure BGZF -- the caller then falls back
// to the sequential gunzip_all.  A truncated or corrupt block ends the stream there (bgzf_read_block fails => sam_read1 < 0).
struct BgzfBlock { size_t in_off, hdr, in_len; uint32_t isize; size_t out_off; };
static bool bgzf_index(const Bytes &in, std::vector<BgzfBlock> &blocks)
{
    size_t p = 0, out = 0;
    while (p < in.size()) {
        if (in.size() - p < 18) break;                                   // truncated tail
        const uint8_t *b = in.data() + p;
        if (b[0] != 0x1f || b[1] != 0x8b || b[2] != 8 || !(b[3] & 4)) return false;
        const size_t xlen = b[10] | ((size_t)b[11] << 8);
        if (in.size() - p < 12 + xlen) break;
        size_t bsize = 0;
        for (size_t x = 0; x + 4 <= xlen;) {
            const uint8_t *f = b + 12 + x; const size_t sl = f[2] | ((size_t)f[3] << 8);
            if (f[0] == 'B' && f[1] == 'C' && sl == 2 && x + 6 <= xlen) bsize = (f[4] | ((size_t)f[5] << 8)) + 1;
            x += 4 + sl;
        }
        if (!bsize) return false;                                        // a gzip member without BC: not BGZF
        if (bsize < 12 + xlen + 8 || in.size() - p < bsize) break;       // truncated block
        BgzfBlock k; k.in_off = p; k.hdr = 12 + xlen; k.in_len = bsize; k.isize = (uint32_t)b[bsize - 4] | ((uint32_t)b[bsize - 3] << 8) | ((uint32_t)b[bsize - 2] << 16) | ((uint32_t)b[bsize - 1] << 24);
        if (k.isize > 65536) break;                                      // a BGZF block inflates to at most 64 KiB: a damaged trailer ends the stream here
        k.out_off = out; out += k.isize;
        blocks.push_back(k);
        p += bsize;
    }
    return true;
}
static bool bgzf_inflate_mt(const Bytes &in, Bytes &out)
{
    std::vector<BgzfBlock> blocks;
    if (!bgzf_index(in, blocks)) return false;
    const size_t total = blocks.empty() ? 0 : blocks.back().out_off + blocks.back().isize;
    out.reserve(total + 1); out.resize(total);
    const size_t GRP = 64, ng = (blocks.size() + GRP - 1) / GRP;
    std::atomic<size_t> first_bad{blocks.size()}; std::atomic<bool> crc_warned{false};
    parallel_for(ng, [&](size_t g) {
        z_stream zs; memset(&zs, 0, sizeof zs);
        if (inflateInit2(&zs, -15) != Z_OK) { size_t b = g * GRP, cur = first_bad.load(); while (b < cur && !first_bad.compare_exchange_weak(cur, b)) {} return; }
        for (size_t i = g * GRP; i < blocks.size() && i < (g + 1) * GRP; ++i) {
            const BgzfBlock &k = blocks[i];
            zs.next_in = (Bytef *)in.data() + k.in_off + k.hdr; zs.avail_in = (uInt)(k.in_len - k.hdr - 8);
            zs.next_out = out.data() + k.out_off; zs.avail_out = k.isize;
            const int rc = k.isize || zs.avail_in ? inflate(&zs, Z_FINISH) : Z_STREAM_END;
            const uint8_t *tr = (const uint8_t *)in.data() + k.in_off + k.in_len - 8;
            const uint32_t want_crc = (uint32_t)tr[0] | ((uint32_t)tr[1] << 8) | ((uint32_t)tr[2] << 16) | ((uint32_t)tr[3] << 24);
            // the reference's htslib (1.3) does not look at the block CRC: a block that inflates to its ISIZE is used as it is.  The
            // same here (the output must not differ), but a mismatch is reported once
            if (rc == Z_STREAM_END && zs.avail_out == 0 && (uint32_t)crc32(crc32(0L, Z_NULL, 0), (const Bytef *)out.data() + k.out_off, k.isize) != want_crc &&
                !crc_warned.exchange(true))
                fprintf(stderr, "[bgzf] CRC mismatch in block %zu (data kept, as htslib 1.3 does)\n", i);
            if (rc != Z_STREAM_END || zs.avail_out != 0) { size_t cur = first_bad.load(); while (i < cur && !first_bad.compare_exchange_weak(cur, i)) {} break; }
            inflateReset(&zs);
        }
        inflateEnd(&zs);
    });
    if (first_bad.load() < blocks.size()) out.resize(blocks[first_bad.load()].out_off);
    return true;
}

static void push_name(Records &r, const char *s, size_t n)
{
    r.names.insert(r.names.end(), s, s + n); r.names.push_back(0);
    r.name_off.push_back((uint64_t)r.names.size());
    r.qhash.push_back(hash_name(s, n));
}

// aux walk: find first NM (bam_aux2i semantics) and first XS (bam_aux2A semantics)
static void scan_aux(const uint8_t *p, const uint8_t *end, int32_t &nm, int8_t &xs, int8_t &nh)
{
    bool got_nm = false, got_xs = false, got_nh = false;
    nm = 0; xs = 0; nh = 0;
    while (p + 3 <= end) {
        const uint8_t *tag = p; uint8_t type = p[2]; p += 3;
        const uint8_t *val = p; size_t sz = 0;
        switch (type) {
        case 'A': case 'c': case 'C': sz = 1; break;
        case 's': case 'S': sz = 2; break;
        case 'i': case 'I': case 'f': sz = 4; break;
        case 'd': sz = 8; break;
        case 'Z': case 'H': { const uint8_t *q = p; while (q < end && *q) ++q; sz = (size_t)(q - p) + 1; break; }
        case 'B': {
            if (p + 5 > end) return;
            uint8_t st = p[0]; uint32_t cnt = rd32(p + 1);
            size_t es = (st == 'c' || st == 'C') ? 1 : (st == 's' || st == 'S') ? 2 : 4;
            sz = 5 + (size_t)cnt * es; break;
        }
        default: return;
        }
        if (p + sz > end) return;
        if (!got_nm && tag[0] == 'N' && tag[1] == 'M') {
            got_nm = true;
            switch (type) {                                            // bam_aux2i, htslib/sam.c:1332-1342
            case 'c': nm = (int8_t)val[0]; break;
            case 'C': nm = val[0]; break;
            case 's': { int16_t v; memcpy(&v, val, 2); nm = v; break; }
            case 'S': { uint16_t v; memcpy(&v, val, 2); nm = v; break; }
            case 'i': nm = rdi32(val); break;
            case 'I': nm = (int32_t)rd32(val); break;
            default: nm = 0;
            }
        }
        if (!got_xs && tag[0] == 'X' && tag[1] == 'S') {
            got_xs = true;
            xs = (type == 'A') ? (int8_t)val[0] : (int8_t)1;           // bam_aux2A returns 0 for non-'A' => is_rev=1 (bam2gtf.c:37)
            if (xs == 0) xs = 1;
        }
        if (!got_nh && tag[0] == 'N' && tag[1] == 'H') {
            got_nh = true;
            long v = 0;
            switch (type) {
            case 'c': v = (int8_t)val[0]; break;
            case 'C': v = val[0]; break;
            case 's': { int16_t x; memcpy(&x, val, 2); v = x; break; }
            case 'S': { uint16_t x; memcpy(&x, val, 2); v = x; break; }
            case 'i': v = rdi32(val); break;
            case 'I': v = (long)rd32(val); break;
            default: v = 0;
            }
            nh = v == 1 ? 1 : 2;
        }
        p += sz;
    }
}

static bool parse_bam(Bytes &d, Header &h, Records &r, std::string &err)
{
    const uint8_t *p = d.data(), *end = p + d.size();
    if (end - p < 12 || memcmp(p, "BAM\1", 4)) { err = "not a BAM stream"; return false; }
    int32_t l_text = rdi32(p + 4); p += 8;
    if (l_text < 0 || (size_t)l_text + 4 > (size_t)(end - p)) { err = "truncated or damaged BAM header"; return false; }
    h.text.assign((const char *)p, (size_t)l_text);
    while (!h.text.empty() && h.text.back() == 0) h.text.pop_back();
    p += l_text;
    int32_t n_ref = rdi32(p); p += 4;
    if (n_ref < 0) { err = "damaged BAM header (n_ref < 0)"; return false; }
    for (int i = 0; i < n_ref; ++i) {
        if ((size_t)(end - p) < 4) { err = "truncated BAM header"; return false; }
        int32_t l_name = rdi32(p); p += 4;
        if (l_name <= 0 || (size_t)l_name + 4 > (size_t)(end - p)) { err = "truncated or damaged BAM header"; return false; }
        std::string name((const char *)p, l_name > 0 ? (size_t)l_name - 1 : 0); p += l_name;
        h.add(name, rd32(p)); p += 4;
    }
    // pass 1 (sequential, touches 36 bytes per record): record boundaries, as sam_read1 would accept them
    std::vector<const uint8_t *> recs;
    while (p + 4 <= end) {
        int32_t bs = rdi32(p);
        if (bs < 32 || p + 4 + bs > end) break;                        // truncated tail: sam_read1 < 0 ends the loop
        const uint8_t *c = p + 4;
        const uint32_t l_qname = rd32(c + 8) & 0xff, n_cigar = rd32(c + 12) & 0xffff; const int32_t l_seq = rdi32(c + 16);
        if (c + 32 + l_qname + 4 * (size_t)n_cigar + ((size_t)l_seq + 1) / 2 + (size_t)l_seq > p + 4 + bs) break;
        recs.push_back(p);
        p += 4 + bs;
    }
    // pass 2: chunks of records sized (names, CIGAR words, raw bytes) in parallel, placed by a prefix sum, filled in parallel
    const size_t n = recs.size(), base = r.n();
    const size_t CH = 16384, nch = (n + CH - 1) / CH;
    std::vector<size_t> c_cig(nch + 1, 0), c_nam(nch + 1, 0), c_raw(nch + 1, 0);
    parallel_for(nch, [&](size_t k) {
        size_t cg = 0, nm = 0, rw = 0;
        for (size_t i = k * CH; i < n && i < (k + 1) * CH; ++i) {
            const uint8_t *c = recs[i] + 4;
            const uint32_t l_qname = rd32(c + 8) & 0xff;
            cg += rd32(c + 12) & 0xffff;
            nm += (l_qname ? strnlen((const char *)c + 32, l_qname) : 0) + 1;
            rw += 4 + (size_t)rdi32(recs[i]);
        }
        c_cig[k + 1] = cg; c_nam[k + 1] = nm; c_raw[k + 1] = rw;
    });
    for (size_t k = 0; k < nch; ++k) { c_cig[k + 1] += c_cig[k]; c_nam[k + 1] += c_nam[k]; c_raw[k + 1] += c_raw[k]; }
    const size_t cig0 = r.cigar.size(), nam0 = r.names.size(), raw0 = r.raw.size();
    r.tid.resize(base + n); r.pos.resize(base + n); r.flag.resize(base + n); r.l_qseq.resize(base + n); r.nm.resize(base + n); r.xs.resize(base + n); r.nh.resize(base + n);
    r.qhash.resize(base + n); r.cigar_off.resize(base + n + 1); r.name_off.resize(base + n + 1);
    r.cigar.resize(cig0 + c_cig[nch]); r.names.resize(nam0 + c_nam[nch]);
    // raw record bodies (for `filter`'s re-emission): the inflated stream itself is adopted when it is the first input
    const bool adopt = r.keep_raw && raw0 == 0 && base == 0;
    if (r.keep_raw) { if (!adopt) r.raw.resize(raw0 + c_raw[nch]); r.raw_off.resize(base + n + 1); if (adopt && n) r.raw_off[0] = (uint64_t)(recs[0] - d.data()); }
    parallel_for(nch, [&](size_t k) {
        size_t cg = cig0 + c_cig[k], nm = nam0 + c_nam[k], rw = raw0 + c_raw[k];
        for (size_t i = k * CH; i < n && i < (k + 1) * CH; ++i) {
            const uint8_t *rp = recs[i], *c = rp + 4; const int32_t bs = rdi32(rp);
            const uint32_t bmn = rd32(c + 8), fnc = rd32(c + 12);
            const int32_t l_seq = rdi32(c + 16);
            const uint32_t l_qname = bmn & 0xff, n_cigar = fnc & 0xffff;
            const uint8_t *q = c + 32, *cig = q + l_qname;
            const uint8_t *aux = cig + 4 * (size_t)n_cigar + ((size_t)l_seq + 1) / 2 + (size_t)l_seq;
            const size_t j = base + i;
            r.tid[j] = rdi32(c); r.pos[j] = rdi32(c + 4); r.flag[j] = (uint16_t)(fnc >> 16); r.l_qseq[j] = l_seq;
            const size_t ln = l_qname ? strnlen((const char *)q, l_qname) : 0;
            memcpy(r.names.data() + nm, q, ln); r.names[nm + ln] = 0; nm += ln + 1;
            r.name_off[j + 1] = (uint64_t)nm; r.qhash[j] = hash_name((const char *)q, ln);
            if (n_cigar) memcpy(r.cigar.data() + cg, cig, 4 * (size_t)n_cigar);
            cg += n_cigar; r.cigar_off[j + 1] = (uint64_t)cg;
            int32_t nmv; int8_t xs, nh; scan_aux(aux, rp + 4 + bs, nmv, xs, nh);
            r.nm[j] = nmv; r.xs[j] = xs; r.nh[j] = nh;
            if (adopt) r.raw_off[j + 1] = (uint64_t)(rp - d.data()) + 4 + (size_t)bs;
            else if (r.keep_raw) { memcpy(r.raw.data() + rw, rp, 4 + (size_t)bs); rw += 4 + (size_t)bs; r.raw_off[j + 1] = rw; }
        }
    });
    if (adopt) r.raw = std::move(d);
    return true;
}

// hts_reg2bin(beg, end, 14, 5)
static int reg2bin(int64_t beg, int64_t end)
{
    int l, s = 14, t = ((1 << (5 * 3)) - 1) / 7;
    for (--end, l = 5; l > 0; --l, s += 3, t -= 1 << (l * 3))
        if (beg >> s == end >> s) return t + (int)(beg >> s);
    return 0;
}

static const char *CIGAR_OPS = "MIDNSHP=XB";

template <class V> static void put32(V &v, uint32_t x) { uint8_t b[4]; memcpy(b, &x, 4); v.insert(v.end(), b, b + 4); }

static uint8_t nt16(char c)
{
    switch (c) {
    case '=': return 0; case 'A': case 'a': return 1; case 'C': case 'c': return 2; case 'M': case 'm': return 3;
    case 'G': case 'g': return 4; case 'R': case 'r': return 5; case 'S': case 's': return 6; case 'V': case 'v': return 7;
    case 'T': case 't': return 8; case 'W': case 'w': return 9; case 'Y': case 'y': return 10; case 'H': case 'h': return 11;
    case 'K': case 'k': return 12; case 'D': case 'd': return 13; case 'B': case 'b': return 14; default: return 15;
    }
}

// one SAM text line -> SoA row (+ optional BAM encoding).  Returns false on a malformed line (sam_parse1 -> -2).
static bool parse_sam_line(char *line, size_t len, const Header &h, Records &r)
{
    char *f[11]; int nf = 0; char *p = line, *lend = line + len;
    f[nf++] = p;
    while (nf < 11) { char *t = (char *)memchr(p, '\t', (size_t)(lend - p)); if (!t) break; *t = 0; p = t + 1; f[nf++] = p; }
    if (nf < 11) return false;
    char *aux = (char *)memchr(f[10], '\t', (size_t)(lend - f[10]));
    if (aux) { *aux = 0; ++aux; }
    char *e;
    long flag = strtol(f[1], &e, 0); if (*e) return false;
    int tid = -1;
    if (strcmp(f[2], "*")) { if (h.names.empty()) return false; tid = h.name2id(f[2]); }
    long pos = strtol(f[3], &e, 10) - 1; if (*e) return false;
    if (pos < 0 && tid >= 0) tid = -1;
    if (tid < 0) flag |= 4;
    long mapq = strtol(f[4], &e, 10); if (*e) return false;
    size_t co = r.cigar.size(); uint32_t n_cigar = 0; int64_t rlen = 0, qlen = 0;
    if (f[5][0] != '*') {
        for (char *q = f[5]; *q;) {
            long l = strtol(q, &q, 10);
            const char *o = *q ? strchr(CIGAR_OPS, *q) : nullptr;
            if (!o) { r.cigar.resize(co); return false; }
            uint32_t op = (uint32_t)(o - CIGAR_OPS);
            r.cigar.push_back(((uint32_t)l << 4) | op); ++n_cigar; ++q;
            if (op == 0 || op == 2 || op == 3 || op == 7 || op == 8) rlen += l;
            if (op == 0 || op == 1 || op == 4 || op == 7 || op == 8) qlen += l;
        }
        if (n_cigar == 0 || n_cigar >= 65536) { r.cigar.resize(co); return false; }
    } else flag |= 4;
    int32_t l_seq = 0;
    if (strcmp(f[9], "*")) {
        l_seq = (int32_t)strlen(f[9]);
        if (n_cigar && qlen != l_seq) { r.cigar.resize(co); return false; }
    }
    if (strcmp(f[10], "*") && (int32_t)strlen(f[10]) != l_seq) { r.cigar.resize(co); return false; }
    // aux
    int32_t nm = 0; int8_t xs = 0, nh = 0; bool got_nm = false, got_xs = false, got_nh = false;
    std::vector<uint8_t> auxenc;
    for (char *a = aux; a && *a;) {
        char *t = strchr(a, '\t'); if (t) *t = 0;
        size_t al = strlen(a);
        if (al < 5) { r.cigar.resize(co); return false; }
        char type = a[3]; char *v = a + 5;
        if (!got_nm && a[0] == 'N' && a[1] == 'M') { got_nm = true; nm = (type == 'i' || type == 'I') ? (int32_t)(*v == '-' ? strtol(v, nullptr, 10) : (long)strtoul(v, nullptr, 10)) : 0; }
        if (!got_xs && a[0] == 'X' && a[1] == 'S') { got_xs = true; xs = (type == 'A' || type == 'a' || type == 'c' || type == 'C') ? (int8_t)*v : (int8_t)1; if (!xs) xs = 1; }
        if (!got_nh && a[0] == 'N' && a[1] == 'H') { got_nh = true; nh = ((type == 'i' || type == 'I') && strtol(v, nullptr, 10) == 1) ? 1 : 2; }
        if (r.keep_raw) {
            auxenc.push_back((uint8_t)a[0]); auxenc.push_back((uint8_t)a[1]);
            if (type == 'A' || type == 'a' || type == 'c' || type == 'C') { auxenc.push_back('A'); auxenc.push_back((uint8_t)*v); }
            else if (type == 'i' || type == 'I') {
                if (*v == '-') {
                    long x = strtol(v, nullptr, 10);
                    if (x >= INT8_MIN) { auxenc.push_back('c'); auxenc.push_back((uint8_t)(int8_t)x); }
                    else if (x >= INT16_MIN) { int16_t y = (int16_t)x; auxenc.push_back('s'); auxenc.insert(auxenc.end(), (uint8_t *)&y, (uint8_t *)&y + 2); }
                    else { int32_t y = (int32_t)x; auxenc.push_back('i'); auxenc.insert(auxenc.end(), (uint8_t *)&y, (uint8_t *)&y + 4); }
                } else {
                    unsigned long x = strtoul(v, nullptr, 10);
                    if (x <= UINT8_MAX) { auxenc.push_back('C'); auxenc.push_back((uint8_t)x); }
                    else if (x <= UINT16_MAX) { uint16_t y = (uint16_t)x; auxenc.push_back('S'); auxenc.insert(auxenc.end(), (uint8_t *)&y, (uint8_t *)&y + 2); }
                    else { uint32_t y = (uint32_t)x; auxenc.push_back('I'); auxenc.insert(auxenc.end(), (uint8_t *)&y, (uint8_t *)&y + 4); }
                }
            } else if (type == 'f') { float x = (float)strtod(v, nullptr); auxenc.push_back('f'); auxenc.insert(auxenc.end(), (uint8_t *)&x, (uint8_t *)&x + 4); }
            else if (type == 'd') { double x = strtod(v, nullptr); auxenc.push_back('d'); auxenc.insert(auxenc.end(), (uint8_t *)&x, (uint8_t *)&x + 8); }
            else if (type == 'Z' || type == 'H') { auxenc.push_back((uint8_t)type); auxenc.insert(auxenc.end(), (uint8_t *)v, (uint8_t *)v + strlen(v) + 1); }
            else if (type == 'B') {
                char st = *v; char *q = v + 1; int32_t cnt = 0;
                for (char *s = q; *s; ++s) if (*s == ',') ++cnt;
                auxenc.push_back('B'); auxenc.push_back((uint8_t)st);
                auxenc.insert(auxenc.end(), (uint8_t *)&cnt, (uint8_t *)&cnt + 4);
                while (*q == ',') {
                    ++q;
                    if (st == 'f') { float x = (float)strtod(q, &q); auxenc.insert(auxenc.end(), (uint8_t *)&x, (uint8_t *)&x + 4); }
                    else {
                        long x = (st == 'c' || st == 's' || st == 'i') ? strtol(q, &q, 0) : (long)strtoul(q, &q, 0);
                        size_t es = (st == 'c' || st == 'C') ? 1 : (st == 's' || st == 'S') ? 2 : 4;
                        auxenc.insert(auxenc.end(), (uint8_t *)&x, (uint8_t *)&x + es);
                    }
                }
            } else { r.cigar.resize(co); return false; }
        }
        if (!t) break;
        a = t + 1;
    }
    size_t lq = strlen(f[0]);
    if (lq > 254) { r.cigar.resize(co); return false; }
    r.tid.push_back(tid); r.pos.push_back((int32_t)pos); r.flag.push_back((uint16_t)flag); r.l_qseq.push_back(l_seq);
    r.nm.push_back(nm); r.xs.push_back(xs); r.nh.push_back(nh);
    push_name(r, f[0], lq);
    r.cigar_off.push_back((uint64_t)r.cigar.size());
    if (r.keep_raw) {
        Bytes &o = r.raw; size_t base = o.size();
        put32(o, 0);                                                   // block_size placeholder
        put32(o, (uint32_t)tid); put32(o, (uint32_t)pos);
        int64_t endpos = pos + ((flag & 4) ? 1 : (n_cigar ? rlen : 1));
        if (!(flag & 4) && n_cigar && rlen == 0) endpos = pos;         // hts_reg2bin(pos, pos+0): same arithmetic as htslib
        uint32_t bin = (uint32_t)reg2bin(pos, endpos);
        put32(o, (bin << 16) | (((uint32_t)mapq & 0xff) << 8) | (uint32_t)(lq + 1));
        put32(o, ((uint32_t)flag << 16) | n_cigar);
        put32(o, (uint32_t)l_seq);
        int32_t mtid = -1;
        if (!strcmp(f[6], "=")) mtid = tid; else if (strcmp(f[6], "*")) mtid = h.name2id(f[6]);
        long mpos = strtol(f[7], nullptr, 10) - 1; if (mpos < 0 && mtid >= 0) mtid = -1;
        put32(o, (uint32_t)mtid); put32(o, (uint32_t)mpos); put32(o, (uint32_t)strtol(f[8], nullptr, 10));
        o.insert(o.end(), (uint8_t *)f[0], (uint8_t *)f[0] + lq + 1);
        for (uint32_t i = 0; i < n_cigar; ++i) put32(o, r.cigar[co + i]);
        size_t sb = o.size(); o.resize(sb + ((size_t)l_seq + 1) / 2, 0);
        for (int32_t i = 0; i < l_seq; ++i) o[sb + (i >> 1)] |= (uint8_t)(nt16(f[9][i]) << ((~i & 1) << 2));
        if (strcmp(f[10], "*")) for (int32_t i = 0; i < l_seq; ++i) o.push_back((uint8_t)(f[10][i] - 33));
        else o.insert(o.end(), (size_t)l_seq, (uint8_t)0xff);
        o.insert(o.end(), auxenc.begin(), auxenc.end());
        uint32_t bs = (uint32_t)(o.size() - base - 4); memcpy(o.data() + base, &bs, 4);
        r.raw_off.push_back(o.size());
    }
    return true;
}

static void sam_header_line(const char *p, size_t len, size_t raw_len, Header &h)
{
    h.text.append(p, raw_len); h.text.push_back('\n');
    if (len > 3 && !memcmp(p, "@SQ", 3)) {
        std::string line(p, len), sn; uint32_t ln = 0;
        size_t s = 0;
        while (s < line.size()) {
            size_t t = line.find('\t', s); if (t == std::string::npos) t = line.size();
            if (t - s > 3 && !line.compare(s, 3, "SN:")) sn = line.substr(s + 3, t - s - 3);
            else if (t - s > 3 && !line.compare(s, 3, "LN:")) ln = (uint32_t)strtoul(line.c_str() + s + 3, nullptr, 10);
            s = t + 1;
        }
        if (!sn.empty()) h.add(sn, ln);
    }
}

// Lines of [p, end) parsed in place (the byte behind every line is overwritten with NUL; end[0] must be writable).
// hdr != nullptr: '@' lines are header lines wherever they stand (sequential mode).  Returns false at a malformed record.
static bool parse_sam_range(char *p, char *end, const Header &h, Header *hdr, Records &r)
{
    while (p < end) {
        char *nl = (char *)memchr(p, '\n', (size_t)(end - p));
        char *le = nl ? nl : end;
        size_t len = (size_t)(le - p);
        if (len && p[len - 1] == '\r') --len;
        if (len && p[0] == '@' && hdr) sam_header_line(p, len, (size_t)(le - p), *hdr);
        else if (len) {
            p[len] = 0;
            if (!parse_sam_line(p, len, h, r)) return false;
        }
        if (!nl) break;
        p = nl + 1;
    }
    return true;
}

static void sam_malformed(size_t row) { fprintf(stderr, "[lr2rmats_b200] malformed SAM record at row %zu; input truncated here (sam_read1 < 0)\n", row); }

static bool parse_sam(Bytes &d, Header &h, Records &r, std::string &err)
{
    (void)err;
    IoTrace tr;
    d.push_back(0);                                                    // writable byte behind the last line
    char *p = (char *)d.data(), *end = p + d.size() - 1;
    // the header: leading '@' (and empty) lines, sequential
    while (p < end) {
        char *nl = (char *)memchr(p, '\n', (size_t)(end - p));
        char *le = nl ? nl : end;
        size_t len = (size_t)(le - p);
        if (len && p[len - 1] == '\r') --len;
        if (len && p[0] != '@') break;
        if (len) sam_header_line(p, len, (size_t)(le - p), h);
        if (!nl) { p = end; break; }
        p = nl + 1;
    }
    const size_t body = (size_t)(end - p);
    const int nt = host_threads();
    bool at_inside = false;                                            // header lines inside the body?  (scanned by all threads)
    if (nt > 1 && body >= (1u << 20)) {
        static const char at[2] = {'\n', '@'};
        const size_t SC = 16u << 20, nsc = (body + SC - 1) / SC;
        std::atomic<bool> hit{false};
        parallel_for(nsc, [&](size_t k) { const size_t o = k * SC, e = o + SC + 1 < body ? o + SC + 1 : body; if (memmem(p + o, e - o, at, 2)) hit = true; });
        at_inside = hit;
    }
    if (nt <= 1 || body < (1u << 20) || at_inside) {                   // small, single thread, or header lines inside the body
        if (!parse_sam_range(p, end, h, &h, r)) sam_malformed(r.n());
        return true;
    }
    // the body: chunks cut at line ends, parsed into chunk-local batches in parallel, concatenated in order
    size_t nch = (size_t)nt * 4; if (nch > body >> 18) nch = body >> 18; if (nch < 1) nch = 1;
    std::vector<char *> cut(nch + 1, end);
    cut[0] = p;
    for (size_t k = 1; k < nch; ++k) {
        char *q = p + body / nch * k; if (q < cut[k - 1]) q = cut[k - 1];
        char *nl = (char *)memchr(q, '\n', (size_t)(end - q));
        cut[k] = nl ? nl + 1 : end;
    }
    tr.lap(" sam header");
    std::vector<Records> part(nch);
    std::vector<uint8_t> ok(nch, 1);
    parallel_for(nch, [&](size_t k) {
        part[k].keep_raw = r.keep_raw;
        ok[k] = parse_sam_range(cut[k], cut[k + 1], h, nullptr, part[k]);
    });
    tr.lap(" sam chunks");
    size_t used = nch;
    for (size_t k = 0; k < nch; ++k) if (!ok[k]) { used = k + 1; break; }
    std::vector<size_t> b_rec(used + 1, r.n()), b_cig(used + 1, r.cigar.size()), b_nam(used + 1, r.names.size()), b_raw(used + 1, r.raw.size());
    for (size_t k = 0; k < used; ++k) {
        b_rec[k + 1] = b_rec[k] + part[k].n(); b_cig[k + 1] = b_cig[k] + part[k].cigar.size();
        b_nam[k + 1] = b_nam[k] + part[k].names.size(); b_raw[k + 1] = b_raw[k] + part[k].raw.size();
    }
    const size_t nn = b_rec[used];
    r.tid.resize(nn); r.pos.resize(nn); r.flag.resize(nn); r.l_qseq.resize(nn); r.nm.resize(nn); r.xs.resize(nn); r.nh.resize(nn); r.qhash.resize(nn);
    r.cigar_off.resize(nn + 1); r.name_off.resize(nn + 1); r.cigar.resize(b_cig[used]); r.names.resize(b_nam[used]);
    if (r.keep_raw) { r.raw.resize(b_raw[used]); r.raw_off.resize(nn + 1); }
    parallel_for(used, [&](size_t k) {
        const Records &q = part[k]; const size_t m = q.n(), o = b_rec[k];
        if (!m) return;
        memcpy(r.tid.data() + o, q.tid.data(), m * 4); memcpy(r.pos.data() + o, q.pos.data(), m * 4); memcpy(r.flag.data() + o, q.flag.data(), m * 2);
        memcpy(r.l_qseq.data() + o, q.l_qseq.data(), m * 4); memcpy(r.nm.data() + o, q.nm.data(), m * 4); memcpy(r.xs.data() + o, q.xs.data(), m); memcpy(r.nh.data() + o, q.nh.data(), m);
        memcpy(r.qhash.data() + o, q.qhash.data(), m * 8);
        memcpy(r.cigar.data() + b_cig[k], q.cigar.data(), q.cigar.size() * 4); memcpy(r.names.data() + b_nam[k], q.names.data(), q.names.size());
        for (size_t i = 0; i < m; ++i) { r.cigar_off[o + i + 1] = (uint64_t)(q.cigar_off[i + 1] + b_cig[k]); r.name_off[o + i + 1] = (uint64_t)(q.name_off[i + 1] + b_nam[k]); }
        if (r.keep_raw) {
            memcpy(r.raw.data() + b_raw[k], q.raw.data(), q.raw.size());
            for (size_t i = 0; i < m; ++i) r.raw_off[o + i + 1] = q.raw_off[i + 1] + b_raw[k];
        }
    });
    tr.lap(" sam merge");
    if (!ok[used - 1]) sam_malformed(r.n());
    return true;
}

bool read_alignments(const std::string &path, Header &h, Records &r, std::string &err)
{
    IoTrace tr;
    Bytes raw;
    if (!slurp(path, raw, err)) return false;
    tr.lap("read file");
    bool ok;
    if (raw.size() >= 2 && raw[0] == 0x1f && raw[1] == 0x8b) {
        Bytes d;
        if (!bgzf_inflate_mt(raw, d) && !gunzip_all(raw, d, err)) return false;
        raw.clear(); raw.shrink_to_fit();
        tr.lap("inflate");
        ok = (d.size() >= 4 && !memcmp(d.data(), "BAM\1", 4)) ? parse_bam(d, h, r, err) : parse_sam(d, h, r, err);
    } else ok = parse_sam(raw, h, r, err);
    tr.lap("parse");
    return ok;
}

// ---- BGZF writer (64 KiB blocks, BC extra field, EOF marker), bgzf.c block format
// one block: src[0..n) -> dst (>= n + 1024 bytes); returns the block's length
static size_t bgzf_block(uint8_t *buf, size_t cap, const uint8_t *src, size_t n)
{
    z_stream zs; memset(&zs, 0, sizeof zs);
    deflateInit2(&zs, Z_DEFAULT_COMPRESSION, Z_DEFLATED, -15, 8, Z_DEFAULT_STRATEGY);
    zs.next_in = (Bytef *)src; zs.avail_in = (uInt)n;
    zs.next_out = buf + 18; zs.avail_out = (uInt)(cap - 18 - 8);
    deflate(&zs, Z_FINISH);
    size_t clen = zs.total_out; deflateEnd(&zs);
    static const uint8_t hdr[12] = {0x1f, 0x8b, 8, 4, 0, 0, 0, 0, 0, 0xff, 6, 0};
    memcpy(buf, hdr, 12); buf[12] = 'B'; buf[13] = 'C'; buf[14] = 2; buf[15] = 0;
    uint16_t bsize = (uint16_t)(clen + 25); memcpy(buf + 16, &bsize, 2);
    uint32_t crc = (uint32_t)crc32(crc32(0L, nullptr, 0), src, (uInt)n), isz = (uint32_t)n;
    memcpy(buf + 18 + clen, &crc, 4); memcpy(buf + 22 + clen, &isz, 4);
    return clen + 26;
}

bool write_bam(FILE *out, const Header &h, const Records &r, const uint32_t *idx, int64_t n, std::string &err)
{
    if (!r.keep_raw) { err = "records were read without raw BAM bodies"; return false; }
    const size_t BLK = 0xff00, CAP = 0x10000 + 1024;
    std::vector<uint8_t> s;
    s.insert(s.end(), {'B', 'A', 'M', 1});
    put32(s, (uint32_t)h.text.size()); s.insert(s.end(), h.text.begin(), h.text.end());
    put32(s, (uint32_t)h.names.size());
    for (size_t i = 0; i < h.names.size(); ++i) {
        put32(s, (uint32_t)h.names[i].size() + 1);
        s.insert(s.end(), h.names[i].begin(), h.names[i].end()); s.push_back(0);
        put32(s, h.lens[i]);
    }
    {   // header in its own block(s), like bam_hdr_write + flush
        std::vector<uint8_t> buf(CAP);
        for (size_t o = 0; o < s.size(); o += BLK) { size_t k = s.size() - o < BLK ? s.size() - o : BLK; fwrite(buf.data(), 1, bgzf_block(buf.data(), CAP, s.data() + o, k), out); }
    }
    // the selected records form one byte stream cut into BLK-byte blocks; groups of blocks are gathered and deflated in
    // parallel (a block finds its first record by binary search on the running record sizes) and written in order
    std::vector<uint64_t> pre((size_t)n + 1, 0);
    for (int64_t k = 0; k < n; ++k) pre[(size_t)k + 1] = pre[(size_t)k] + (r.raw_off[idx[k] + 1] - r.raw_off[idx[k]]);
    const uint64_t total = pre[(size_t)n];
    const size_t nblk = (size_t)((total + BLK - 1) / BLK), GRP = (size_t)host_threads() * 8;
    std::vector<uint8_t> comp(GRP * CAP), plain(GRP * BLK);
    std::vector<size_t> clen(GRP);
    for (size_t b0 = 0; b0 < nblk; b0 += GRP) {
        const size_t nb = nblk - b0 < GRP ? nblk - b0 : GRP;
        parallel_for(nb, [&](size_t j) {
            const uint64_t lo = (uint64_t)(b0 + j) * BLK, hi = lo + BLK < total ? lo + BLK : total;
            uint8_t *dst = plain.data() + j * BLK;
            size_t k = (size_t)(std::upper_bound(pre.begin(), pre.end(), lo) - pre.begin()) - 1;    // record holding byte lo
            for (uint64_t at = lo; at < hi; ++k) {
                const uint64_t r0 = r.raw_off[idx[k]], skip = at - pre[k], len = pre[k + 1] - pre[k];
                const uint64_t take = len - skip < hi - at ? len - skip : hi - at;
                memcpy(dst + (at - lo), r.raw.data() + r0 + skip, (size_t)take);
                at += take;
            }
            clen[j] = bgzf_block(comp.data() + j * CAP, CAP, dst, (size_t)(hi - lo));
        });
        for (size_t j = 0; j < nb; ++j) fwrite(comp.data() + j * CAP, 1, clen[j], out);
    }
    static const uint8_t eof[28] = {0x1f, 0x8b, 8, 4, 0, 0, 0, 0, 0, 0xff, 6, 0, 'B', 'C', 2, 0, 0x1b, 0, 3, 0, 0, 0, 0, 0, 0, 0, 0, 0};
    fwrite(eof, 1, 28, out);
    return true;
}

}  // namespace lrb
