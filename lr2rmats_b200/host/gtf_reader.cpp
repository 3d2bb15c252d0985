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

bool read_gtf(const std::string &fn, const Header &h, Anno &a, bool gtf_mode, std::string &err)
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

bool read_sj(const std::string &fn, ChrNames &cn, SjTable &sj, std::string &err)
{
    FILE *fp = fopen(fn.c_str(), "r");
    if (!fp) { err = "Can not open splice-junction file \"" + fn + "\""; return false; }
    static char line[1024], ref[1024]; ref[0] = 0;
    struct Row { int tid, don, acc, uniq, multi; };
    std::vector<Row> rows;
    int don = 0, acc = 0, strand = 0, motif = 0, anno = 0, uniq = 0, multi = 0, over = 0;
    while (fgets(line, 1024, fp) != NULL) {
        sscanf(line, "%s\t%d\t%d\t%d\t%d\t%d\t%d\t%d\t%d", ref, &don, &acc, &strand, &motif, &anno, &uniq, &multi, &over);
        rows.push_back({cn.get_id(ref), don, acc, uniq, multi});
    }
    fclose(fp);
    std::stable_sort(rows.begin(), rows.end(), [](const Row &x, const Row &y) {
        if (x.tid != y.tid) return x.tid < y.tid;
        if (x.don != y.don) return x.don < y.don;
        return x.acc < y.acc; });                                      // sj_group_comp, gtf.c:414-420
    for (auto &r : rows) { sj.tid.push_back(r.tid); sj.don.push_back(r.don); sj.acc.push_back(r.acc); sj.uniq.push_back(r.uniq); sj.multi.push_back(r.multi); }
    return true;
}

}  // namespace lrb
