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
#include <zlib.h>
#include "lrb_host.h"

namespace lrb {

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

static bool slurp(const std::string &path, std::vector<uint8_t> &buf, std::string &err)
{
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

// concatenated gzip members (BGZF blocks are gzip members, bgzf.c) -> one buffer
static bool gunzip_all(const std::vector<uint8_t> &in, std::vector<uint8_t> &out, std::string &err)
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

static void push_name(Records &r, const char *s, size_t n)
{
    r.names.insert(r.names.end(), s, s + n); r.names.push_back(0);
    r.name_off.push_back((uint32_t)r.names.size());
    r.qhash.push_back(hash_name(s, n));
}

// aux walk: find first NM (bam_aux2i semantics) and first XS (bam_aux2A semantics)
static void scan_aux(const uint8_t *p, const uint8_t *end, int32_t &nm, int8_t &xs)
{
    bool got_nm = false, got_xs = false;
    nm = 0; xs = 0;
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
        p += sz;
    }
}

static bool parse_bam(const std::vector<uint8_t> &d, Header &h, Records &r, std::string &err)
{
    const uint8_t *p = d.data(), *end = p + d.size();
    if (end - p < 12 || memcmp(p, "BAM\1", 4)) { err = "not a BAM stream"; return false; }
    int32_t l_text = rdi32(p + 4); p += 8;
    if (p + l_text + 4 > end) { err = "truncated BAM header"; return false; }
    h.text.assign((const char *)p, (size_t)l_text);
    while (!h.text.empty() && h.text.back() == 0) h.text.pop_back();
    p += l_text;
    int32_t n_ref = rdi32(p); p += 4;
    for (int i = 0; i < n_ref; ++i) {
        if (p + 4 > end) { err = "truncated BAM header"; return false; }
        int32_t l_name = rdi32(p); p += 4;
        if (p + l_name + 4 > end) { err = "truncated BAM header"; return false; }
        std::string name((const char *)p, l_name > 0 ? (size_t)l_name - 1 : 0); p += l_name;
        h.add(name, rd32(p)); p += 4;
    }
    while (p + 4 <= end) {
        int32_t bs = rdi32(p);
        if (bs < 32 || p + 4 + bs > end) break;                        // truncated tail: sam_read1 < 0 ends the loop
        const uint8_t *c = p + 4;
        int32_t tid = rdi32(c), pos = rdi32(c + 4);
        uint32_t bmn = rd32(c + 8), fnc = rd32(c + 12);
        int32_t l_seq = rdi32(c + 16);
        uint32_t l_qname = bmn & 0xff, n_cigar = fnc & 0xffff;
        const uint8_t *q = c + 32;
        const uint8_t *cig = q + l_qname;
        const uint8_t *aux = cig + 4 * (size_t)n_cigar + ((size_t)l_seq + 1) / 2 + (size_t)l_seq;
        if (aux > p + 4 + bs) break;
        r.tid.push_back(tid); r.pos.push_back(pos); r.flag.push_back((uint16_t)(fnc >> 16)); r.l_qseq.push_back(l_seq);
        push_name(r, (const char *)q, l_qname ? strnlen((const char *)q, l_qname) : 0);
        size_t co = r.cigar.size(); r.cigar.resize(co + n_cigar);
        if (n_cigar) memcpy(r.cigar.data() + co, cig, 4 * (size_t)n_cigar);
        r.cigar_off.push_back((uint32_t)r.cigar.size());
        int32_t nm; int8_t xs; scan_aux(aux, p + 4 + bs, nm, xs);
        r.nm.push_back(nm); r.xs.push_back(xs);
        if (r.keep_raw) { r.raw.insert(r.raw.end(), p, p + 4 + bs); r.raw_off.push_back(r.raw.size()); }
        p += 4 + bs;
    }
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

static void put32(std::vector<uint8_t> &v, uint32_t x) { uint8_t b[4]; memcpy(b, &x, 4); v.insert(v.end(), b, b + 4); }

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
    int32_t nm = 0; int8_t xs = 0; bool got_nm = false, got_xs = false;
    std::vector<uint8_t> auxenc;
    for (char *a = aux; a && *a;) {
        char *t = strchr(a, '\t'); if (t) *t = 0;
        size_t al = strlen(a);
        if (al < 5) { r.cigar.resize(co); return false; }
        char type = a[3]; char *v = a + 5;
        if (!got_nm && a[0] == 'N' && a[1] == 'M') { got_nm = true; nm = (type == 'i' || type == 'I') ? (int32_t)(*v == '-' ? strtol(v, nullptr, 10) : (long)strtoul(v, nullptr, 10)) : 0; }
        if (!got_xs && a[0] == 'X' && a[1] == 'S') { got_xs = true; xs = (type == 'A' || type == 'a' || type == 'c' || type == 'C') ? (int8_t)*v : (int8_t)1; if (!xs) xs = 1; }
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
    r.nm.push_back(nm); r.xs.push_back(xs);
    push_name(r, f[0], lq);
    r.cigar_off.push_back((uint32_t)r.cigar.size());
    if (r.keep_raw) {
        std::vector<uint8_t> &o = r.raw; size_t base = o.size();
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

static bool parse_sam(std::vector<uint8_t> &d, Header &h, Records &r, std::string &err)
{
    char *p = (char *)d.data(), *end = p + d.size();
    while (p < end) {
        char *nl = (char *)memchr(p, '\n', (size_t)(end - p));
        char *le = nl ? nl : end;
        size_t len = (size_t)(le - p);
        if (len && p[len - 1] == '\r') --len;
        if (len && p[0] == '@') {
            h.text.append(p, (size_t)(le - p)); h.text.push_back('\n');
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
        } else if (len) {
            std::vector<char> tmp(p, p + len); tmp.push_back(0);
            if (!parse_sam_line(tmp.data(), len, h, r)) {
                fprintf(stderr, "[lr2rmats_b200] malformed SAM record at row %zu; input truncated here (sam_read1 < 0)\n", r.n());
                break;
            }
        }
        if (!nl) break;
        p = nl + 1;
    }
    (void)err;
    return true;
}

bool read_alignments(const std::string &path, Header &h, Records &r, std::string &err)
{
    std::vector<uint8_t> raw;
    if (!slurp(path, raw, err)) return false;
    if (raw.size() >= 2 && raw[0] == 0x1f && raw[1] == 0x8b) {
        std::vector<uint8_t> d;
        if (!gunzip_all(raw, d, err)) return false;
        raw.clear(); raw.shrink_to_fit();
        if (d.size() >= 4 && !memcmp(d.data(), "BAM\1", 4)) return parse_bam(d, h, r, err);
        return parse_sam(d, h, r, err);
    }
    return parse_sam(raw, h, r, err);
}

// ---- BGZF writer (64 KiB blocks, BC extra field, EOF marker), bgzf.c block format
static void bgzf_block(FILE *out, const uint8_t *src, size_t n)
{
    uint8_t buf[0x10000 + 1024];
    z_stream zs; memset(&zs, 0, sizeof zs);
    deflateInit2(&zs, Z_DEFAULT_COMPRESSION, Z_DEFLATED, -15, 8, Z_DEFAULT_STRATEGY);
    zs.next_in = (Bytef *)src; zs.avail_in = (uInt)n;
    zs.next_out = buf + 18; zs.avail_out = sizeof buf - 18 - 8;
    deflate(&zs, Z_FINISH);
    size_t clen = zs.total_out; deflateEnd(&zs);
    static const uint8_t hdr[12] = {0x1f, 0x8b, 8, 4, 0, 0, 0, 0, 0, 0xff, 6, 0};
    memcpy(buf, hdr, 12); buf[12] = 'B'; buf[13] = 'C'; buf[14] = 2; buf[15] = 0;
    uint16_t bsize = (uint16_t)(clen + 25); memcpy(buf + 16, &bsize, 2);
    uint32_t crc = (uint32_t)crc32(crc32(0L, nullptr, 0), src, (uInt)n), isz = (uint32_t)n;
    memcpy(buf + 18 + clen, &crc, 4); memcpy(buf + 22 + clen, &isz, 4);
    fwrite(buf, 1, clen + 26, out);
}

bool write_bam(FILE *out, const Header &h, const Records &r, const uint32_t *idx, int64_t n, std::string &err)
{
    if (!r.keep_raw) { err = "records were read without raw BAM bodies"; return false; }
    std::vector<uint8_t> s;
    s.insert(s.end(), {'B', 'A', 'M', 1});
    put32(s, (uint32_t)h.text.size()); s.insert(s.end(), h.text.begin(), h.text.end());
    put32(s, (uint32_t)h.names.size());
    for (size_t i = 0; i < h.names.size(); ++i) {
        put32(s, (uint32_t)h.names[i].size() + 1);
        s.insert(s.end(), h.names[i].begin(), h.names[i].end()); s.push_back(0);
        put32(s, h.lens[i]);
    }
    const size_t BLK = 0xff00;
    auto flush = [&](bool all) {
        size_t o = 0;
        while (s.size() - o >= BLK || (all && o < s.size())) { size_t k = s.size() - o < BLK ? s.size() - o : BLK; bgzf_block(out, s.data() + o, k); o += k; }
        s.erase(s.begin(), s.begin() + (long)o);
    };
    flush(true);                                                       // header in its own block(s), like bam_hdr_write + flush
    for (int64_t k = 0; k < n; ++k) {
        uint32_t i = idx[k];
        s.insert(s.end(), r.raw.begin() + (long)r.raw_off[i], r.raw.begin() + (long)r.raw_off[i + 1]);
        if (s.size() >= 4 * BLK) flush(false);
    }
    flush(true);
    static const uint8_t eof[28] = {0x1f, 0x8b, 8, 4, 0, 0, 0, 0, 0, 0xff, 6, 0, 'B', 'C', 2, 0, 0x1b, 0, 3, 0, 0, 0, 0, 0, 0, 0, 0, 0};
    fwrite(eof, 1, 28, out);
    return true;
}

}  // namespace lrb
