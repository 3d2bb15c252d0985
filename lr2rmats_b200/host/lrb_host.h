// lrb_host.h -- host side of lr2rmats_b200 above the C ABI (include/lr2rmats_b200.h).
//
// What lives here is what the reference does around its per-alignment hot path:
// decode alignments into structure-of-arrays batches (the reference uses htslib's
// sam_read1; we decode SAM text / BAM ourselves straight into SoA, never building
// bam1_t), read the annotation GTF and STAR SJ.out.tab with the reference's exact
// quirks (gtf.c:431-521), and print GTF / BED / detail / summary byte-compatibly
// (gtf.c:597-632, update_gtf.c:297-419,535-576).
#ifndef LRB_HOST_H
#define LRB_HOST_H

#include <cstdint>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <functional>
#include <memory>
#include <string>
#include <unordered_map>
#include <vector>
#include "../../include/lr2rmats_b200.h"

namespace lrb {

// ---------------------------------------------------------------- alignments
struct Header {
    std::vector<std::string> names;
    std::vector<uint32_t> lens;
    std::string text;                              // SAM header text (for BAM re-emission)
    std::unordered_map<std::string, int> index;
    int name2id(const std::string &s) const { auto it = index.find(s); return it == index.end() ? -1 : it->second; }
    void add(const std::string &n, uint32_t len) { if (!index.count(n)) index[n] = (int)names.size(); names.push_back(n); lens.push_back(len); }
};

// byte buffer whose resize() leaves new bytes uninitialised: the large I/O buffers are first touched by the worker
// threads that fill them (page faults in parallel) instead of being zero-filled by one thread
template <class T> struct NoInitAlloc : std::allocator<T> {
    template <class U> struct rebind { using other = NoInitAlloc<U>; };
    template <class U> void construct(U *p) noexcept { ::new ((void *)p) U; }
    template <class U, class... A> void construct(U *p, A &&...a) { ::new ((void *)p) U(std::forward<A>(a)...); }
};
using Bytes = std::vector<uint8_t, NoInitAlloc<uint8_t>>;

struct Records {
    std::vector<int32_t> tid, pos, l_qseq, nm;
    std::vector<uint16_t> flag;
    std::vector<int8_t> xs;
    std::vector<int8_t> nh;                        // NH:i tag (bam2sj's bam_is_uniq_NH, parse_bam.c:239-247): 0 absent, 1 NH == 1, 2 otherwise
    std::vector<uint64_t> qhash;
    std::vector<uint64_t> cigar_off{0}; std::vector<uint32_t> cigar;   // 64-bit running totals: no silent wrap past 2^32 ops / name bytes
    std::vector<uint64_t> name_off{0};
    std::vector<char> names;                       // NUL-terminated qnames
    bool keep_raw = false;                         // keep BAM-encoded records for `filter` re-emission
    std::vector<uint64_t> raw_off{0};
    Bytes raw;                                     // block_size-prefixed BAM records
    size_t n() const { return tid.size(); }
    const char *qname(size_t i) const { return names.data() + name_off[i]; }
    lrb_batch view() const;
};

uint64_t hash_name(const char *s, size_t n);

// host threads used by the decoders / encoders / emitters (LRB_THREADS, default: every hardware thread) and the loop helper
int host_threads();
void parallel_for(size_t n, const std::function<void(size_t)> &fn);
// LRB_IO_TRACE=1: stage timings of the host side (decode, table readers, engine call, emitters) on stderr
struct IoTrace {
    bool on = getenv("LRB_IO_TRACE") != nullptr; std::chrono::steady_clock::time_point t = std::chrono::steady_clock::now();
    void lap(const char *what) { if (!on) return; auto n = std::chrono::steady_clock::now(); fprintf(stderr, "[lrb io] %-16s %.3f s\n", what, std::chrono::duration<double>(n - t).count()); t = n; }
};

// Reads SAM text or BAM (BGZF) -- autodetected like sam_open(..., "rb").  Returns false and sets err on failure.
bool read_alignments(const std::string &path, Header &h, Records &r, std::string &err);
// Writes a BAM (BGZF) stream with the header and the selected raw records (filter's stdout, bam_filter.c:127-159).
bool write_bam(FILE *out, const Header &h, const Records &r, const uint32_t *idx, int64_t n, std::string &err);

// ---------------------------------------------------------------- annotation
struct Anno {                                       // read_anno_trans / read_gtf_trans result
    std::vector<int32_t> tid, start, end, gene;
    std::vector<uint8_t> is_rev;
    std::vector<uint32_t> exon_off{0};
    std::vector<int32_t> es, ee;
    std::vector<std::string> gene_id, gene_name, trans_id, trans_name;
    int gene_n = 0;                                 // T->gene_n (gtf.c:495 / :553)
    size_t n() const { return tid.size(); }
    lrb_anno view() const;
    lrb_chains chains() const;
};
// gtf_mode=false: read_anno_trans (gtf.c:468); true: read_gtf_trans (gtf.c:524).  Fatal format errors -> false + err.
bool read_gtf(const std::string &fn, const Header &h, Anno &a, bool gtf_mode, std::string &err);

struct ChrNames {                                   // chr_name_t, gtf.c:336-412
    std::vector<std::string> names;
    int get_id(const std::string &s) { for (size_t i = 0; i < names.size(); ++i) if (names[i] == s) return (int)i; names.push_back(s); return (int)names.size() - 1; }
    void seed(const Header &h) { for (auto &s : h.names) get_id(s); }
};
struct SjTable {
    std::vector<int32_t> tid, don, acc, uniq, multi;
    lrb_sj view() const;
};
bool read_sj(const std::string &fn, ChrNames &cn, SjTable &sj, std::string &err);   // read_sj_group, gtf.c:431

// ------------------------------------------------------------------ emitters
struct RowNames {                                   // where trans_id / trans_name of a read row come from
    const Records *rec = nullptr;                   // BAM mode: qname of read_idx[row]
    const Anno *chains = nullptr;                   // -m g mode: names from the input GTF
};

void emit_bam2gtf(FILE *out, const lrb_exon_result &ex, const Records &rec, const ChrNames &cn, const char *src);
void emit_update_outputs(const lrb_update_result &res, const RowNames &rn, const Anno &anno, const Header &h, const ChrNames &cn,
                         const char *src, int anno_gene_n, int anno_trans_n,
                         FILE *updated, FILE *bam_gtf, FILE *detail, FILE *known, FILE *novel, FILE *unrecog, FILE *summary, FILE *bed);
// updated GTF / summary / BED straight from the self-contained table (lrb_update_fetch_table): same bytes as
// emit_update_outputs(updated, summary, bed) without the per-read tables
void emit_update_table(const lrb_trans_table &tab, const lrb_bed_list *bed, const int32_t *summary_counts, const RowNames &rn, const Anno &anno,
                       const Header &h, const ChrNames &cn, const char *src, int anno_gene_n, int anno_trans_n, FILE *updated, FILE *summary, FILE *bed_fp);
void emit_unique(FILE *out, const lrb_unique_result &res, const RowNames &rn, const ChrNames &cn, const char *src, bool intersect);

// -------------------------------------------------------------------- engine
// The CLI core is engine-agnostic so that the oracle tool (oracle/port_main.cpp) can drive the CPU restatement through
// the very same readers and emitters.  The product binary binds this to the CUDA C ABI only (no CPU path exists there).
struct Engine {
    void *self = nullptr;
    int (*set_tables)(void *, const lrb_anno *anno, const lrb_anno *rm, const lrb_sj *sj) = nullptr;
    int (*filter)(void *, const lrb_batch *, const lrb_filter_params *, lrb_filter_result *) = nullptr;
    int (*bam2gtf)(void *, const lrb_batch *, const lrb_exon_params *, lrb_exon_result *) = nullptr;
    int (*update)(void *, const lrb_batch *, const lrb_chains *, const lrb_exon_params *, const lrb_update_params *, lrb_update_result *) = nullptr;
    int (*unique)(void *, const lrb_batch *, const lrb_chains *, const lrb_exon_params *, const lrb_update_params *, lrb_unique_result *) = nullptr;
    // optional: run update and return only what -o / -y / -E print (NULL: the CLI uses `update` and the full tables)
    int (*update_table)(void *, const lrb_batch *, const lrb_chains *, const lrb_exon_params *, const lrb_update_params *,
                        lrb_trans_table *, lrb_bed_list *, int32_t *summary) = nullptr;
    int (*bam2sj)(void *, const lrb_batch *, const uint8_t *is_uniq, const lrb_sj_params *, lrb_sj *) = nullptr;   // bam2sj_core, parse_bam.c:896
    // stable sort of n records by (k0, k1, k2): the `sort -n` of src/sort_gtf.sh; *perm = record indices in sorted order
    int (*sort3)(void *, const uint32_t *k0, const uint32_t *k1, const uint32_t *k2, int64_t n, const uint32_t **perm) = nullptr;
    const char *(*error)(void *) = nullptr;
};

int cli_main(int argc, char **argv, Engine &eng);   // main.c:37-49 dispatch + the four subcommands

}  // namespace lrb
#endif
