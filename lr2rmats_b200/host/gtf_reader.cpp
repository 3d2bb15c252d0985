// gtf_reader.cpp -- annotation GTF and STAR SJ.out.tab readers.
//
// Output parity depends on the reference's parsing quirks, so the same libc primitives are used
// with the same buffer lifetimes: fgets() into a 1024-byte line (longer lines split, Q7), one sscanf() per line whose
// unmatched fields keep their previous values, attribute lookup by first substring hit (gtf.c:317-326), a new transcript
// whenever transcript_id changes (gtf.c:496), exons sorted by (start,end) and transcript tid/strand/start/end re-derived
// from the first/last sorted exon (gtf.c:94-100), gene_n counted per exon line against the last *started* transcript's
// gene (gtf.c:495; read_gtf_trans compares gene_name instead, gtf.c:553 -- Q8).
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include "lrb_host.h"

namespace lrb {

lrb_anno Anno::view() const
{
    lrb_anno a;
    a.n_trans = (int32_t)n(); a.n_exon = (int64_t)es.size();
    a.tid = tid.data(); a.start = start.data(); a.end = end.data(); a.is_rev = is_rev.data(); a.gene = gene.data();
    a.exon_off = exon_off.data(); a.exon_start = es.data(); a.exon_end = ee.data();
    return a;
}
lrb_chains Anno::chains() const
{
    lrb_chains c;
    c.n = (int64_t)n(); c.tid = tid.data(); c.is_rev = is_rev.data(); c.exon_off = exon_off.data();
    c.exon_start = es.data(); c.exon_end = ee.data();
    return c;
}
lrb_sj SjTable::view() const
{
    lrb_sj s; s.n = (int64_t)tid.size(); s.tid = tid.data(); s.don = don.data(); s.acc = acc.data(); s.uniq_c = uniq.data(); s.multi_c = multi.data();
    return s;
}

// gtf_add_info, gtf.c:317-326
static void attr_value(const char *info, const char *tag, char *out)
{
    size_t tl = strlen(tag);
    for (size_t i = 0; info[i]; ++i)
        if (strncmp(info + i, tag, tl) == 0) { sscanf(info + i + tl + 2, "%[^\"]", out); return; }
}

struct PendingTrans {
    std::vector<int32_t> tid, s, e; std::vector<uint8_t> rev;
    std::string tname, tidn, gname, gid;
};

static bool flush_trans(PendingTrans &t, Anno &a, std::unordered_map<std::string, int> &genes, std::string &err)
{
    size_t n = t.s.size();
    std::vector<size_t> ord(n);
    for (size_t i = 0; i < n; ++i) ord[i] = i;
    for (size_t i = 1; i < n; ++i)
        if (t.rev[i] != t.rev[0]) { err = "Strands of exons do NOT match."; return false; }   // trans_exon_comp, gtf.c:40-42
    std::stable_sort(ord.begin(), ord.end(), [&](size_t x, size_t y) { return t.s[x] != t.s[y] ? t.s[x] < t.s[y] : t.e[x] < t.e[y]; });
    a.tid.push_back(t.tid[ord[0]]); a.is_rev.push_back(t.rev[ord[0]]);
    a.start.push_back(t.s[ord[0]]); a.end.push_back(t.e[ord[n - 1]]);
    for (size_t i = 0; i < n; ++i) { a.es.push_back(t.s[ord[i]]); a.ee.push_back(t.e[ord[i]]); }
    a.exon_off.push_back((uint32_t)a.es.size());
    a.gene_id.push_back(t.gid); a.gene_name.push_back(t.gname); a.trans_id.push_back(t.tidn); a.trans_name.push_back(t.tname);
    auto it = genes.find(t.gid);
    int g = it == genes.end() ? (genes[t.gid] = (int)genes.size()) : it->second;
    a.gene.push_back(g);
    t.tid.clear(); t.s.clear(); t.e.clear(); t.rev.clear();
    return true;
}

// the reference's reader, statement for statement: one thread, libc parsing with all its quirks
static bool read_gtf_sequential(const std::string &fn, const Header &h, Anno &a, bool gtf_mode, std::string &err)
{
    FILE *fp = fopen(fn.c_str(), "r");
    if (!fp) { err = "fail to open file '" + fn + "'"; return false; }
    static char line[1024], ref[1024], type[1024], add_info[1024], gname[1024], gid[1024], trans_name[1024], trans_id[1024];
    ref[0] = type[0] = add_info[0] = 0; int start = 0, end = 0; char strand = 0;
    memset(gname, 0, sizeof gname); memset(gid, 0, sizeof gid); memset(trans_name, 0, sizeof trans_name); memset(trans_id, 0, sizeof trans_id);
    std::string last_tid, last_gid;
    PendingTrans t;
    std::unordered_map<std::string, int> genes;
    bool ok = true;
    while (fgets(line, 1024, fp) != NULL) {
        if (line[0] == '#') continue;
        sscanf(line, "%s\t%*s\t%s\t%d\t%d\t%*s\t%c\t%*s\t%[^\n]", ref, type, &start, &end, &strand, add_info);
        if (strcmp(type, "exon") != 0) continue;
        uint8_t is_rev = (strand == '-' ? 1 : 0);
        int tid = h.name2id(ref);
        memset(gid, 0, strlen(gid)); attr_value(add_info, "gene_id", gid);
        memset(gname, 0, strlen(gname)); attr_value(add_info, "gene_name", gname);
        if (!gid[0] && !gname[0]) { err = "GTF format error in " + fn + ". (No gene id or gene name found."; ok = false; break; }
        if (!gid[0]) strcpy(gid, gname); else if (!gname[0]) strcpy(gname, gid);
        memset(trans_id, 0, strlen(trans_id)); attr_value(add_info, "transcript_id", trans_id);
        memset(trans_name, 0, strlen(trans_name)); attr_value(add_info, "transcript_name", trans_name);
        if (!trans_id[0] && !trans_name[0]) { err = "GTF format error in " + fn + ". (No transcript id or transcript name found."; ok = false; break; }
        if (!trans_id[0]) strcpy(trans_id, trans_name); else if (!trans_name[0]) strcpy(trans_name, trans_id);

        a.gene_n += (last_gid != (gtf_mode ? gname : gid)) ? 1 : 0;
        if (last_tid != trans_id) {
            if (!t.s.empty() && !flush_trans(t, a, genes, err)) { ok = false; break; }
            t.tname = trans_name; t.tidn = trans_id; t.gname = gname; t.gid = gid;
            last_tid = trans_id; last_gid = gtf_mode ? gname : gid;
        }
        t.tid.push_back(tid); t.s.push_back(start); t.e.push_back(end); t.rev.push_back(is_rev);
    }
    if (ok && !t.s.empty() && !flush_trans(t, a, genes, err)) ok = false;
    fclose(fp);
    return ok;
}

// ---- all-threads reader for well-formed files -------------------------------------------------------------------------
// A line is "plain" when the reference's fgets + sscanf see exactly its tab-separated columns: shorter than the 1024-byte
// line buffer, at least 9 tab-separated columns, columns 1..8 non-empty and free of blanks, columns 4 and 5 plain decimal
// numbers, column 7 one character, column 9 non-empty.  On plain lines no parser state survives from one line to the next
// (every sscanf conversion succeeds, the attribute buffers are cleared per line), so lines can be parsed independently: the
// file is read once, cut into chunks at line ends, the chunks are parsed by all host threads, and one ordered pass groups
// the exon lines into transcripts exactly like the sequential reader.  A single line that is not plain (or any error) sends
// the WHOLE file through read_gtf_sequential -- quirks and error texts then come from the very same code as before.
namespace {
struct Sv { const char *p; uint32_t n; bool eq(const Sv &o) const { return n == o.n && !memcmp(p, o.p, n); } };
struct ExonLine { int32_t tid, start, end; uint8_t rev; Sv gid, gname, tidn, tname; };

// gtf_add_info (gtf.c:317-326): first substring hit of the tag, value = the characters two behind the tag up to the next quote
bool plain_attr(const char *info, size_t len, const char *tag, size_t tl, Sv &out)
{
    out.p = info; out.n = 0;
    const char *hit = (const char *)memmem(info, len, tag, tl);
    if (!hit) return true;
    const char *v = hit + tl + 2, *end = info + len;
    if (v > end) return false;                                   // the reference would read behind the line: leave it to that code
    const char *q = (const char *)memchr(v, '"', (size_t)(end - v));
    out.p = v; out.n = (uint32_t)((q ? q : end) - v);
    return true;
}
bool plain_int(const char *p, const char *e, int32_t &v)
{
    if (p == e || e - p > 9) return false;
    int32_t x = 0;
    for (; p < e; ++p) { if (*p < '0' || *p > '9') return false; x = x * 10 + (*p - '0'); }
    v = x; return true;
}
// parses the lines of [p, end); false when a line is not plain
bool parse_plain_chunk(const char *p, const char *end, const Header &h, std::vector<ExonLine> &out)
{
    std::string last_ref; int last_tid = -1; bool have_ref = false;
    while (p < end) {
        const char *nl = (const char *)memchr(p, '\n', (size_t)(end - p));
        const char *le = nl ? nl : end;
        const size_t raw = (size_t)(le - p) + (nl ? 1 : 0);
        if (raw >= 1023) return false;                           // fgets(line, 1024) would split it (Q7)
        if (p < le && p[0] == '#') { p = le + 1; continue; }
        const char *c[10]; int nc = 0; c[nc++] = p;
        for (const char *q = p; nc < 9 && (q = (const char *)memchr(q, '\t', (size_t)(le - q))); ) c[nc++] = ++q;
        if (nc < 9) return false;
        c[9] = le + 1;
        for (int k = 0; k < 8; ++k) {
            const char *b = c[k], *e = c[k + 1] - 1;
            if (b >= e) return false;
            for (const char *q = b; q < e; ++q) if (*q == ' ' || *q == '\r' || *q == '\v' || *q == '\f') return false;
        }
        const char *info = c[8]; size_t il = (size_t)(le - info);
        if (il == 0 || info[0] == ' ' || info[0] == '\t' || info[0] == '\r' || info[0] == '\v' || info[0] == '\f' || memchr(info, 0, il)) return false;
        if (c[7] - 1 - c[6] != 1) return false;
        ExonLine x;
        if (!plain_int(c[3], c[4] - 1, x.start) || !plain_int(c[4], c[5] - 1, x.end)) return false;
        if (c[3] - 1 - c[2] == 4 && !memcmp(c[2], "exon", 4)) {
            x.rev = c[6][0] == '-';
            const size_t rl = (size_t)(c[1] - 1 - c[0]);
            if (!have_ref || last_ref.size() != rl || memcmp(last_ref.data(), c[0], rl)) { last_ref.assign(c[0], rl); last_tid = h.name2id(last_ref); have_ref = true; }
            x.tid = last_tid;
            if (!plain_attr(info, il, "gene_id", 7, x.gid) || !plain_attr(info, il, "gene_name", 9, x.gname) ||
                !plain_attr(info, il, "transcript_id", 13, x.tidn) || !plain_attr(info, il, "transcript_name", 15, x.tname)) return false;
            if ((!x.gid.n && !x.gname.n) || (!x.tidn.n && !x.tname.n)) return false;        // format error: reported by the sequential code
            if (!x.gid.n) x.gid = x.gname; else if (!x.gname.n) x.gname = x.gid;
            if (!x.tidn.n) x.tidn = x.tname; else if (!x.tname.n) x.tname = x.tidn;
            out.push_back(x);
        }
        p = le + 1;
    }
    return true;
}
}  // namespace

bool read_file_bytes(const std::string &path, Bytes &buf);       // align_reader.cpp

bool read_gtf(const std::string &fn, const Header &h, Anno &a, bool gtf_mode, std::string &err)
{
    Bytes buf;
    if (getenv("LRB_GTF_SEQUENTIAL") || !read_file_bytes(fn, buf) || buf.size() < (1u << 16)) return read_gtf_sequential(fn, h, a, gtf_mode, err);
    const char *p = (const char *)buf.data(), *end = p + buf.size();
    size_t nch = (size_t)host_threads() * 4; if (nch > buf.size() >> 16) nch = buf.size() >> 16; if (nch < 1) nch = 1;
    std::vector<const char *> cut(nch + 1, end); cut[0] = p;
    for (size_t k = 1; k < nch; ++k) {
        const char *q = p + buf.size() / nch * k; if (q < cut[k - 1]) q = cut[k - 1];
        const char *nl = (const char *)memchr(q, '\n', (size_t)(end - q));
        cut[k] = nl ? nl + 1 : end;
    }
    std::vector<std::vector<ExonLine>> part(nch);
    std::vector<uint8_t> ok(nch, 1);
    parallel_for(nch, [&](size_t k) { part[k].reserve((size_t)(cut[k + 1] - cut[k]) / 160 + 16); ok[k] = parse_plain_chunk(cut[k], cut[k + 1], h, part[k]); });
    for (size_t k = 0; k < nch; ++k) if (!ok[k]) return read_gtf_sequential(fn, h, a, gtf_mode, err);
    // ordered grouping pass: gene_n (gtf.c:495 / :553), a new transcript whenever transcript_id changes (gtf.c:496)
    PendingTrans t; std::unordered_map<std::string, int> genes;
    Sv last_tid{"", 0}, last_gid{"", 0};
    for (size_t k = 0; k < nch; ++k)
        for (const ExonLine &x : part[k]) {
            const Sv &g = gtf_mode ? x.gname : x.gid;
            a.gene_n += last_gid.eq(g) ? 0 : 1;
            if (!last_tid.eq(x.tidn)) {
                if (!t.s.empty() && !flush_trans(t, a, genes, err)) { Anno fresh; a = fresh; return read_gtf_sequential(fn, h, a, gtf_mode, err); }
                t.tname.assign(x.tname.p, x.tname.n); t.tidn.assign(x.tidn.p, x.tidn.n); t.gname.assign(x.gname.p, x.gname.n); t.gid.assign(x.gid.p, x.gid.n);
                last_tid = x.tidn; last_gid = g;
            }
            t.tid.push_back(x.tid); t.s.push_back(x.start); t.e.push_back(x.end); t.rev.push_back(x.rev);
        }
    if (!t.s.empty() && !flush_trans(t, a, genes, err)) { Anno fresh; a = fresh; return read_gtf_sequential(fn, h, a, gtf_mode, err); }
    return true;
}

namespace {
struct SjRow { int tid, don, acc, uniq, multi; };
struct SjLine { Sv ref; int don, acc, uniq, multi; };
void sj_finish(std::vector<SjRow> &rows, SjTable &sj)
{
    std::stable_sort(rows.begin(), rows.end(), [](const SjRow &x, const SjRow &y) {
        if (x.tid != y.tid) return x.tid < y.tid;
        if (x.don != y.don) return x.don < y.don;
        return x.acc < y.acc; });                                      // sj_group_comp, gtf.c:414-420
    const size_t n = rows.size();
    sj.tid.resize(n); sj.don.resize(n); sj.acc.resize(n); sj.uniq.resize(n); sj.multi.resize(n);
    for (size_t i = 0; i < n; ++i) { sj.tid[i] = rows[i].tid; sj.don[i] = rows[i].don; sj.acc[i] = rows[i].acc; sj.uniq[i] = rows[i].uniq; sj.multi[i] = rows[i].multi; }
}
// plain SJ.out.tab lines (9 tab-separated columns, column 1 a short name, columns 2..9 plain decimal numbers): what the
// reference's sscanf reads is exactly the columns, no state carries over.  false: some line is not plain
bool parse_plain_sj_chunk(const char *p, const char *end, std::vector<SjLine> &out)
{
    while (p < end) {
        const char *nl = (const char *)memchr(p, '\n', (size_t)(end - p));
        const char *le = nl ? nl : end;
        if ((size_t)(le - p) + (nl ? 1 : 0) >= 1023) return false;
        const char *c[10]; int nc = 0; c[nc++] = p;
        for (const char *q = p; nc < 9 && (q = (const char *)memchr(q, '\t', (size_t)(le - q))); ) c[nc++] = ++q;
        if (nc < 9) return false;
        const char *e9 = (const char *)memchr(c[8], '\t', (size_t)(le - c[8]));
        c[9] = (e9 ? e9 : (le > c[8] && le[-1] == '\r' ? le - 1 : le)) + 1;
        const size_t rl = (size_t)(c[1] - 1 - c[0]);
        if (rl == 0 || rl >= 100) return false;
        for (const char *q = c[0]; q < c[1] - 1; ++q) if (*q == ' ' || *q == '\r' || *q == '\v' || *q == '\f' || !*q) return false;
        int32_t v[8];
        for (int k = 1; k < 9; ++k) if (!plain_int(c[k], c[k + 1] - 1, v[k - 1])) return false;
        out.push_back({Sv{c[0], (uint32_t)rl}, v[0], v[1], v[5], v[6]});
        p = le + 1;
    }
    return true;
}
}  // namespace

// the reference's reader (read_sj_group, gtf.c:431-449), one thread
static bool read_sj_sequential(const std::string &fn, ChrNames &cn, SjTable &sj, std::string &err)
{
    FILE *fp = fopen(fn.c_str(), "r");
    if (!fp) { err = "Can not open splice-junction file \"" + fn + "\""; return false; }
    static char line[1024], ref[1024]; ref[0] = 0;
    std::vector<SjRow> rows;
    int don = 0, acc = 0, strand = 0, motif = 0, anno = 0, uniq = 0, multi = 0, over = 0;
    while (fgets(line, 1024, fp) != NULL) {
        sscanf(line, "%s\t%d\t%d\t%d\t%d\t%d\t%d\t%d\t%d", ref, &don, &acc, &strand, &motif, &anno, &uniq, &multi, &over);
        rows.push_back({cn.get_id(ref), don, acc, uniq, multi});
    }
    fclose(fp);
    sj_finish(rows, sj);
    return true;
}

bool read_sj(const std::string &fn, ChrNames &cn, SjTable &sj, std::string &err)
{
    Bytes buf;
    if (getenv("LRB_GTF_SEQUENTIAL") || !read_file_bytes(fn, buf) || buf.size() < (1u << 16)) return read_sj_sequential(fn, cn, sj, err);
    const char *p = (const char *)buf.data(), *end = p + buf.size();
    size_t nch = (size_t)host_threads() * 4; if (nch > buf.size() >> 16) nch = buf.size() >> 16; if (nch < 1) nch = 1;
    std::vector<const char *> cut(nch + 1, end); cut[0] = p;
    for (size_t k = 1; k < nch; ++k) {
        const char *q = p + buf.size() / nch * k; if (q < cut[k - 1]) q = cut[k - 1];
        const char *nl = (const char *)memchr(q, '\n', (size_t)(end - q));
        cut[k] = nl ? nl + 1 : end;
    }
    std::vector<std::vector<SjLine>> part(nch);
    std::vector<uint8_t> ok(nch, 1);
    parallel_for(nch, [&](size_t k) { part[k].reserve((size_t)(cut[k + 1] - cut[k]) / 24 + 16); ok[k] = parse_plain_sj_chunk(cut[k], cut[k + 1], part[k]); });
    for (size_t k = 0; k < nch; ++k) if (!ok[k]) return read_sj_sequential(fn, cn, sj, err);
    // chromosome ids in order of first appearance (get_chr_id, gtf.c:389-403): one ordered pass, the previous name cached
    std::vector<SjRow> rows; size_t total = 0;
    for (auto &v : part) total += v.size();
    rows.reserve(total);
    Sv last{"", 0}; int last_id = -1;
    for (size_t k = 0; k < nch; ++k)
        for (const SjLine &x : part[k]) {
            if (last_id < 0 || !last.eq(x.ref)) { last = x.ref; last_id = cn.get_id(std::string(x.ref.p, x.ref.n)); }
            rows.push_back({last_id, x.don, x.acc, x.uniq, x.multi});
        }
    sj_finish(rows, sj);
    return true;
}

}  // namespace lrb
