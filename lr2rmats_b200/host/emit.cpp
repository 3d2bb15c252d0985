// emit.cpp -- text emitters, byte-compatible with the reference's printers:
//   print_trans (gtf.c:597-605), print_read_trans (gtf.c:607-632), print_bam_detail_trans (update_gtf.c:297-419),
//   the summary.txt / novel_exon.bed blocks of print_trans_summary (update_gtf.c:535-576).
// They consume the SoA result structs of the C ABI plus the name tables kept on the host.
#include <cstring>
#include <string>
#include <vector>
#include "lrb_host.h"

namespace lrb {

namespace {

struct Buf {                                        // append buffer in front of FILE* (fprintf per line is the slow part)
    FILE *fp; std::string s;                        // fp == nullptr: collect only (a chunk formatted by a worker thread)
    explicit Buf(FILE *f) : fp(f) { s.reserve(1 << 20); }
    ~Buf() { flush(); }
    void flush() { if (fp && !s.empty()) { fwrite(s.data(), 1, s.size(), fp); s.clear(); } }
    void put(const char *p) { s.append(p); if (fp && s.size() > (1 << 20) - 4096) flush(); }
    void put(const std::string &p) { s.append(p); if (fp && s.size() > (1 << 20) - 4096) flush(); }
    void ch(char c) { s.push_back(c); }
    void num(long v)                                // "%ld"
    {
        char t[24]; int k = 24;
        unsigned long u = v < 0 ? 0ul - (unsigned long)v : (unsigned long)v;
        do { t[--k] = (char)('0' + u % 10); u /= 10; } while (u);
        if (v < 0) t[--k] = '-';
        s.append(t + k, (size_t)(24 - k));
    }
};

// fn(o, i) for i in [0, n), output in order: the items are formatted by all host threads, a chunk of items per task into
// its own buffer, and the buffers of a group of chunks are written out in sequence
template <class F> void emit_items(FILE *fp, int64_t n, F &&fn)
{
    if (!fp) return;
    const int nt = host_threads();
    const int64_t CH = 1024;
    if (nt <= 1 || n < 4 * CH) { Buf o(fp); for (int64_t i = 0; i < n; ++i) fn(o, i); return; }
    const int64_t nch = (n + CH - 1) / CH, G = (int64_t)nt * 8;
    std::vector<Buf> part; part.reserve((size_t)G);
    for (int64_t j = 0; j < G; ++j) part.emplace_back(nullptr);
    for (int64_t g0 = 0; g0 < nch; g0 += G) {
        const int64_t nb = nch - g0 < G ? nch - g0 : G;
        parallel_for((size_t)nb, [&](size_t j) {
            Buf &o = part[j]; o.s.clear();
            const int64_t lo = (g0 + (int64_t)j) * CH, hi = lo + CH < n ? lo + CH : n;
            for (int64_t i = lo; i < hi; ++i) fn(o, i);
        });
        for (int64_t j = 0; j < nb; ++j) fwrite(part[(size_t)j].s.data(), 1, part[(size_t)j].s.size(), fp);
    }
}

struct TransText { const char *gene_id, *gene_name, *trans_id, *trans_name; };

// one transcript + its exon lines, print_read_trans (gtf.c:612-629)
void put_read_trans(Buf &o, const char *tchr, const char *src, int tstart, int tend, int trev, const TransText &nm, int cov,
                    int n, const int32_t *es, const int32_t *ee, int fs, int le, const char *echr, int erev)
{
    std::string attr;
    if (nm.gene_id[0]) { attr += " gene_id \""; attr += nm.gene_id; attr += "\";"; }
    if (nm.trans_id[0]) { attr += " transcript_id \""; attr += nm.trans_id; attr += "\";"; }
    if (nm.gene_name[0]) { attr += " gene_name \""; attr += nm.gene_name; attr += "\";"; }
    if (nm.trans_name[0]) { attr += " transcript_name \""; attr += nm.trans_name; attr += "\";"; }
    const char *ea = attr.empty() ? "" : attr.c_str() + 1;
    o.put(tchr); o.ch('\t'); o.put(src); o.put("\ttranscript\t"); o.num(tstart); o.ch('\t'); o.num(tend);
    o.put("\t.\t"); o.ch("+-"[trev]); o.put("\t.\t"); o.put(ea);
    o.put(attr.empty() ? "transcript_cov \"" : " transcript_cov \""); o.num(cov); o.put("\";\n");
    for (int k = 0; k < n; ++k) {
        int j = trev ? n - 1 - k : k;
        int s = j == 0 ? fs : es[j], e = j == n - 1 ? le : ee[j];
        o.put(echr); o.ch('\t'); o.put(src); o.put("\texon\t"); o.num(s); o.ch('\t'); o.num(e);
        o.put("\t.\t"); o.ch("+-"[erev]); o.put("\t.\t"); o.put(ea); o.ch('\n');
    }
}

struct Namer {
    const RowNames &rn; const lrb_exon_result &ex;
    Namer(const RowNames &r, const lrb_exon_result &e) : rn(r), ex(e) {}
    const char *tid_(int64_t row) const { return rn.chains ? rn.chains->trans_id[row].c_str() : rn.rec->qname(ex.read_idx ? ex.read_idx[row] : row); }
    const char *tname_(int64_t row) const { return rn.chains ? rn.chains->trans_name[row].c_str() : tid_(row); }
};

}  // namespace

namespace {
void put_summary(FILE *summary, const int32_t *s, int anno_gene_n, int anno_trans_n)   // update_gtf.c:535-570
{
    fprintf(summary, "==== Annotaion ====\n");
    fprintf(summary, "%s\t%d\n", "Genes_of_annotation_GTF", anno_gene_n);
    fprintf(summary, "%s\t%d\n", "Transcripts_of_annotation_GTF", anno_trans_n);
    fprintf(summary, "\n===================\n");
    fprintf(summary, "\n==== Updated information ====\n");
    fprintf(summary, "%s\t%d\n", "Updated_Genes", s[LRB_S_UPD_GENES]);
    fprintf(summary, "%s\t%d\n", "Added_Novel_Transcripts", s[LRB_S_NOVEL_FULL] + s[LRB_S_NOVEL_PARTIAL]);
    fprintf(summary, "%s\t%d\n", "Added_Novel_Full-read_Transcripts", s[LRB_S_NOVEL_FULL]);
    fprintf(summary, "%s\t%d\n", "Added_Novel_Partial-read_Transcripts", s[LRB_S_NOVEL_PARTIAL]);
    fprintf(summary, "%s\t%d\n", "Added_Novel_Exons", s[LRB_S_NOVEL_EXONS]);
    fprintf(summary, "%s\t%d\n", "Added_Novel_Sites", s[LRB_S_NOVEL_SITES]);
    fprintf(summary, "%s\t%d\n", "Added_Novel_Splice_Junctions", s[LRB_S_NOVEL_JUNC]);
    fprintf(summary, "\n=============================\n");
    fprintf(summary, "\n==== Known information ====\n");
    fprintf(summary, "%s\t%d\n", "Known_Transcripts_from_BAM", s[LRB_S_KNOWN_TRANS]);
    fprintf(summary, "%s\t%d\n", "Genes_of_Known_Transcripts_from_BAM", s[LRB_S_KNOWN_GENES]);
    fprintf(summary, "%s\t%d\n", "Uniq_Known_Transcripts_from_BAM", s[LRB_S_UNIQ_KNOWN]);
    fprintf(summary, "\n===========================\n");
    fprintf(summary, "\n==== Novel information ====\n");
    fprintf(summary, "%s\t%d\n", "Novel_Transcript_from_BAM", s[LRB_S_NOVEL_RELIABLE] + s[LRB_S_NOVEL_UNRELIABLE]);
    fprintf(summary, "%s\t%d\n", "Novel_Transcript_from_BAM_with_All_Reliable_Junction", s[LRB_S_NOVEL_RELIABLE]);
    fprintf(summary, "%s\t%d\n", "Uniq_Novel_Transcript_from_BAM_with_All_Reliable_Junction", s[LRB_S_UNIQ_RELIABLE]);
    fprintf(summary, "%s\t%d\n", "Novel_Transcript_from_BAM_with_Unreliable_Junction", s[LRB_S_NOVEL_UNRELIABLE]);
    fprintf(summary, "%s\t%d\n", "Uniq_Novel_Transcript_from_BAM_with_Unreliable_Junction", s[LRB_S_UNIQ_UNRELIABLE]);
    fprintf(summary, "\n===========================\n");
    fprintf(summary, "\n==== Unrecognized information ====\n");
    fprintf(summary, "%s\t%d\n", "Unrecognized_Transcript_from_BAM", s[LRB_S_UNRECOG]);
    fprintf(summary, "%s\t%d\n", "Uniq_Unrecognized_Transcript_from_BAM", s[LRB_S_UNIQ_UNRECOG]);
    fprintf(summary, "\n==================================\n");
}
void put_bed(FILE *bed, const lrb_bed_list &b, const Header &h)                          // update_gtf.c:571-576 (uses the BAM header names)
{
    emit_items(bed, b.n, [&](Buf &o, int64_t i) {
        o.put(h.names[b.tid[i]]); o.ch('\t'); o.num(b.start[i] - 1); o.ch('\t'); o.num(b.end[i]); o.ch('\t');
        o.ch("TIS"[b.type[i]]); o.put("_exon\t"); o.num(b.score[i]); o.ch('\t'); o.ch("+-"[b.is_rev[i]]); o.ch('\n');
    });
}
}  // namespace

// bam2gtf's loop: print_trans for every mapped record (bam2gtf.c:150-156, gtf.c:597-605)
void emit_bam2gtf(FILE *out, const lrb_exon_result &ex, const Records &rec, const ChrNames &cn, const char *src)
{
    emit_items(out, ex.n_reads, [&](Buf &o, int64_t r) {
        uint32_t lo = ex.exon_off[r], hi = ex.exon_off[r + 1];
        if (hi == lo) return;
        const char *q = rec.qname(ex.read_idx ? ex.read_idx[r] : r);
        const char *chr = cn.names[ex.tid[r]].c_str();
        char strand = "+-"[ex.is_rev[r]];
        auto line = [&](const char *feat, int s, int e) {
            o.put(chr); o.ch('\t'); o.put(src); o.ch('\t'); o.put(feat); o.ch('\t'); o.num(s); o.ch('\t'); o.num(e);
            o.put("\t.\t"); o.ch(strand); o.put("\t.\tgene_id \""); o.put(q); o.put("\"; transcript_id \""); o.put(q); o.put("\";\n");
        };
        line("transcript", ex.exon_start[lo], ex.exon_end[hi - 1]);
        for (uint32_t j = lo; j < hi; ++j) line("exon", ex.exon_start[j], ex.exon_end[j]);
    });
}

void emit_update_outputs(const lrb_update_result &res, const RowNames &rn, const Anno &anno, const Header &h, const ChrNames &cn,
                         const char *src, int anno_gene_n, int anno_trans_n,
                         FILE *updated, FILE *bam_gtf, FILE *detail, FILE *known, FILE *novel, FILE *unrecog, FILE *summary, FILE *bed)
{
    const lrb_exon_result &ex = res.ex;
    Namer nm(rn, ex);
    auto gene_of = [&](int64_t row, const char *&gid, const char *&gname) {
        int ref = res.ref_anno[row];
        if (ref >= 0) { gid = anno.gene_id[ref].c_str(); gname = anno.gene_name[ref].c_str(); } else { gid = "NA"; gname = "NA"; }
    };
    auto put_row = [&](Buf &o, int64_t row) {                        // a whole bam_T entry (cov 1)
        uint32_t lo = ex.exon_off[row]; int n = (int)(ex.exon_off[row + 1] - lo);
        TransText t; gene_of(row, t.gene_id, t.gene_name); t.trans_id = nm.tid_(row); t.trans_name = nm.tname_(row);
        const char *chr = cn.names[ex.tid[row]].c_str();
        put_read_trans(o, chr, src, ex.exon_start[lo], ex.exon_end[lo + n - 1], ex.is_rev[row], t, 1, n, ex.exon_start + lo, ex.exon_end + lo,
                       ex.exon_start[lo], ex.exon_end[lo + n - 1], chr, ex.is_rev[row]);
    };
    auto put_list_row = [&](Buf &o, int64_t c, int cov, int ttid, int tstart, int tend, int fs, int le, bool use_merged) {
        int64_t row = res.novel.read[c]; uint32_t lo = ex.exon_off[row] + res.novel.exon_lo[c]; int n = (int)res.novel.exon_n[c];
        int piece = res.novel.piece[c];
        TransText t; gene_of(row, t.gene_id, t.gene_name);
        std::string id = nm.tid_(row), name = nm.tname_(row);
        if (piece >= 0) { id += ".split." + std::to_string(piece); name += ".split." + std::to_string(piece); }
        t.trans_id = id.c_str(); t.trans_name = name.c_str();
        const char *echr = cn.names[ex.tid[row]].c_str();
        int trev = piece >= 0 ? 0 : ex.is_rev[row];
        if (!use_merged) {
            fs = ex.exon_start[lo]; le = ex.exon_end[lo + n - 1];
            if (piece >= 0) { ttid = 0; tstart = 0; tend = 0; } else { ttid = ex.tid[row]; tstart = fs; tend = le; }
        }
        put_read_trans(o, cn.names[ttid].c_str(), src, tstart, tend, trev, t, cov, n, ex.exon_start + lo, ex.exon_end + lo, fs, le, echr, ex.is_rev[row]);
    };

    emit_items(updated, res.updated.n, [&](Buf &o, int64_t i) {
        put_list_row(o, res.updated.cand[i], res.updated.cov[i], res.updated.t_tid[i], res.updated.t_start[i], res.updated.t_end[i],
                     res.updated.first_start[i], res.updated.last_end[i], true);
    });
    emit_items(bam_gtf, ex.n_reads, [&](Buf &o, int64_t r) { put_row(o, r); });
    if (detail) {                                                    // print_bam_detail_trans, update_gtf.c:297-419
        { Buf o(detail);
        o.put("ReadName\tchr\tstrand\tNovel\tGeneID\tGeneName\tExonCount\tExonStart\tExonEnd\tNovelExonCount\tNovelExonIndex\tNovelSiteCount\tNovelSiteIndex\tNovelJunctionCount\tNovelJunctionIndex\tUnreliableJunctionCount\tUnreliableJunctionIndex\n"); }
        emit_items(detail, ex.n_reads, [&](Buf &o, int64_t r) {
            uint32_t lo = ex.exon_off[r]; int n = (int)(ex.exon_off[r + 1] - lo);
            const uint8_t *f = res.exon_flag + lo;
            uint32_t c = res.cls[r];
            int nov = (c & LRB_C_KNOWN) ? 0 : (c & LRB_C_KNOWN_SITE) ? 1 : 2;
            const char *gid, *gname; gene_of(r, gid, gname);
            o.put(nm.tname_(r)); o.ch('\t'); o.put(cn.names[ex.tid[r]]); o.ch('\t'); o.ch("+-"[ex.is_rev[r]]); o.ch('\t'); o.num(nov); o.ch('\t');
            o.put(gid); o.ch('\t'); o.put(gname); o.ch('\t'); o.num(n); o.ch('\t');
            for (int j = 0; j < n; ++j) { if (j) o.ch(','); o.num(ex.exon_start[lo + j]); } o.ch('\t');
            for (int j = 0; j < n; ++j) { if (j) o.ch(','); o.num(ex.exon_end[lo + j]); } o.ch('\t');
            auto idx_list = [&](int cnt_slots, int stride, uint8_t bit0, uint8_t bit1, bool last) {
                // stride 1: one flag per slot (bit0); stride 2: site flags 2j (bit0) and 2j+1 (bit1)
                int cnt = 0;
                for (int j = 0; j < cnt_slots; ++j) { cnt += (f[j] & bit0) ? 1 : 0; if (stride == 2) cnt += (f[j] & bit1) ? 1 : 0; }
                o.num(cnt); o.ch('\t');
                if (cnt == 0) { o.put("NA\t"); return; }
                bool first = true;
                for (int j = 0; j < cnt_slots; ++j) {
                    if (f[j] & bit0) { if (!first) o.ch(','); first = false; o.num(stride == 2 ? 2 * j : j); }
                    if (stride == 2 && (f[j] & bit1)) { if (!first) o.ch(','); first = false; o.num(2 * j + 1); }
                }
                if (!last) o.ch('\t');
            };
            idx_list(n, 1, LRB_F_NOVEL_EXON, 0, false);
            idx_list(n - 1, 2, LRB_F_NOVEL_DON, LRB_F_NOVEL_ACC, false);
            idx_list(n - 1, 1, LRB_F_NOVEL_JUNC, 0, false);
            idx_list(n - 1, 1, LRB_F_UNRELIABLE, 0, true);
            o.ch('\n');
        });
    }
    emit_items(known, res.n_known, [&](Buf &o, int64_t i) { put_row(o, res.known_idx[i]); });
    emit_items(novel, res.novel.n, [&](Buf &o, int64_t c) { put_list_row(o, c, 1, 0, 0, 0, 0, 0, false); });
    emit_items(unrecog, res.n_unrecog, [&](Buf &o, int64_t i) { put_row(o, res.unrecog_idx[i]); });
    if (summary) put_summary(summary, res.summary, anno_gene_n, anno_trans_n);
    if (bed) put_bed(bed, res.bed, h);
}

void emit_update_table(const lrb_trans_table &tab, const lrb_bed_list *bed, const int32_t *summary_counts, const RowNames &rn, const Anno &anno,
                       const Header &h, const ChrNames &cn, const char *src, int anno_gene_n, int anno_trans_n, FILE *updated, FILE *summary, FILE *bed_fp)
{
    {
        emit_items(updated, tab.n, [&](Buf &o, int64_t i) {
            const uint32_t lo = tab.exon_off[i]; const int n = (int)(tab.exon_off[i + 1] - lo);
            const int64_t k = tab.name_idx[i];
            TransText t;
            const int ref = tab.ref_anno[i];
            if (ref >= 0) { t.gene_id = anno.gene_id[ref].c_str(); t.gene_name = anno.gene_name[ref].c_str(); } else { t.gene_id = "NA"; t.gene_name = "NA"; }
            std::string id = rn.chains ? rn.chains->trans_id[k] : std::string(rn.rec->qname(k));
            std::string name = rn.chains ? rn.chains->trans_name[k] : id;
            if (tab.piece[i] >= 0) { id += ".split." + std::to_string(tab.piece[i]); name += ".split." + std::to_string(tab.piece[i]); }
            t.trans_id = id.c_str(); t.trans_name = name.c_str();
            put_read_trans(o, cn.names[tab.t_tid[i]].c_str(), src, tab.t_start[i], tab.t_end[i], tab.t_rev[i], t, tab.cov[i], n, tab.exon_start + lo,
                           tab.exon_end + lo, tab.exon_start[lo], tab.exon_end[lo + n - 1], cn.names[tab.e_tid[i]].c_str(), tab.e_rev[i]);
        });
    }
    if (summary && summary_counts) put_summary(summary, summary_counts, anno_gene_n, anno_trans_n);
    if (bed_fp && bed) put_bed(bed_fp, *bed, h);
}

// unique-gtf: print_read_trans over unique_T or shared_T (unique_gtf.c:147-148)
void emit_unique(FILE *out, const lrb_unique_result &res, const RowNames &rn, const ChrNames &cn, const char *src, bool intersect)
{
    const lrb_exon_result &ex = res.ex;
    Namer nm(rn, ex);
    auto names = [&](int64_t row, TransText &t) {
        t.trans_id = nm.tid_(row); t.trans_name = nm.tname_(row);
        if (rn.chains) { t.gene_id = rn.chains->gene_id[row].c_str(); t.gene_name = rn.chains->gene_name[row].c_str(); }
        else { t.gene_id = t.trans_id; t.gene_name = t.trans_id; }
    };
    if (intersect) {
        emit_items(out, res.n_shared, [&](Buf &o, int64_t i) {
            int64_t row = res.shared_idx[i]; uint32_t lo = ex.exon_off[row]; int n = (int)(ex.exon_off[row + 1] - lo);
            TransText t; names(row, t);
            const char *chr = cn.names[ex.tid[row]].c_str();
            put_read_trans(o, chr, src, ex.exon_start[lo], ex.exon_end[lo + n - 1], ex.is_rev[row], t, 1, n, ex.exon_start + lo, ex.exon_end + lo,
                           ex.exon_start[lo], ex.exon_end[lo + n - 1], chr, ex.is_rev[row]);
        });
    } else {
        emit_items(out, res.uniq.n, [&](Buf &o, int64_t i) {
            int64_t row = res.uniq.cand[i]; uint32_t lo = ex.exon_off[row]; int n = (int)(ex.exon_off[row + 1] - lo);
            TransText t; names(row, t);
            put_read_trans(o, cn.names[res.uniq.t_tid[i]].c_str(), src, res.uniq.t_start[i], res.uniq.t_end[i], ex.is_rev[row], t, res.uniq.cov[i], n,
                           ex.exon_start + lo, ex.exon_end + lo, res.uniq.first_start[i], res.uniq.last_end[i], cn.names[ex.tid[row]].c_str(), ex.is_rev[row]);
        });
    }
}

}  // namespace lrb
