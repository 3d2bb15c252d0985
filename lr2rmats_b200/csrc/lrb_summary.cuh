// lrb_summary.cuh -- argument blocks of the summary-set kernels (lrb_summary.cu)
#pragma once
#include "lrb_kernels.cuh"

namespace lrbk {

// open addressing, 128-bit keys; one 32-byte slot (= one memory sector) carries the key, the minimum stream position and
// the accumulated score of an element, so all atomics on an element stay in one sector.  Empty = all bytes 0xFF
// (the score therefore starts at -1).
struct __align__(32) HashSlot { uint64_t khi, klo, minpos; int32_t score, pad; };
struct HashTab {
    uint64_t cap = 0;                               // number of slots
    HashSlot *slots = nullptr;
};

struct SummaryArgs {
    DRows rows; DExons ex; DTransList list; DMerged upd; int64_t n_upd;
    const uint64_t *n_upd_dev;                      // sum_count_kernel only: the count still lives on the device (n_upd is then an upper bound)
    const int32_t *ref; const int32_t *anno_gene;
    HashTab tab;
    uint32_t *bar_cnt, *bar_seg;                    // [4][n_upd]: inserted tid-0 elements per entry / their exclusive scan
    uint64_t *gene_bar;                             // [n_upd]
    uint32_t *bed_cnt, *bed_off;                    // [n_upd]
    uint32_t *counts;                               // [8]: E D A J G KG partial -
    int32_t *bed_tid, *bed_start, *bed_end, *bed_score; uint8_t *bed_type, *bed_rev;
    int sets;                                       // which sets this run computes: SUM_E | SUM_DAJ | SUM_G | SUM_KG (the merge on the gather root recomputes only what shards cannot
                                                    // add up; 0 = only the probe for split pieces that meet a chain on another chromosome)
    // known reads' genes: first occurrences of (tid, gene) exported for the cross-shard union (optional)
    int2 *kg_pairs;
    // gather root: entry i belongs to shard k iff shard_end[k-1] <= i < shard_end[k] (device array); 0 shards = single-GPU run
    const int64_t *shard_end; int n_shards;
    uint32_t *xs_pl, *xs_hx, *xs_cnt; uint32_t xs_cap;    // probe: piece list, hit list (entries sharing a junction key with a piece of another group), their counters [2] (zeroed by the caller)
    int probe;                                      // 1: the junction keys of split pieces are probed for a meeting with another chromosome / shard (CNT_XLOCUS)
};
enum { SUM_E = 1, SUM_DAJ = 2, SUM_G = 4, SUM_KG = 8, SUM_ALL = 15 };
// counts[] slots of the summary kernels
enum { CNT_E = 0, CNT_D = 1, CNT_A = 2, CNT_J = 3, CNT_G = 4, CNT_KG = 5, CNT_PARTIAL = 6, CNT_XLOCUS = 7 };

void launch_summary_count(const SummaryArgs &a, unsigned long long *n_elems, cudaStream_t st);
// all sets + BED rows; the exon chain runs on st_exon (pass st for a single stream)
void launch_summary_sets(const SummaryArgs &a, const uint32_t *cls, int64_t n_rows, uint64_t *tile_state, uint32_t *ticket, uint64_t *bed_total,
                         cudaStream_t st, cudaStream_t st_exon, cudaEvent_t ev_fork, cudaEvent_t ev_join);

// ---- gather root: the concatenated per-shard tables as one updated_T (lrb_multi.cu)
struct GatheredTable {
    int64_t n = 0, n_exon = 0;
    uint32_t *name_idx = nullptr; int32_t *piece = nullptr, *t_tid = nullptr, *t_start = nullptr, *t_end = nullptr, *e_tid = nullptr, *cov = nullptr, *ref = nullptr;
    uint8_t *t_rev = nullptr, *e_rev = nullptr; uint32_t *exon_off = nullptr;      // n + 1
    int32_t *es = nullptr, *ee = nullptr; uint8_t *flag = nullptr;
};
// view of a gathered table in the shape the summary kernels read (scratch: 5 x n words + n bytes)
void launch_table_view(const GatheredTable &t, uint32_t *ident, uint32_t *zeros, uint32_t *cnt, int32_t *fs, int32_t *le, cudaStream_t st);
// tid-0 elements (split pieces, chromosome 0) of different shards with equal site / junction keys: the per-shard counts do not
// add up then.  shard_end[k] = first entry behind shard k; flag gets 1.
void launch_tid0_coincidence(const SummaryArgs &a, uint32_t *flag, cudaStream_t st);      // shards from a.shard_end / a.n_shards
// distinct (tid, gene) pairs over a pair list: count of distinct keys -> *out
void launch_pairs_distinct(const HashTab &tab, const int2 *pairs, int64_t n, uint32_t *out, cudaStream_t st);

// from lrb_update.cu
void launch_rows_as_list(const DRows &rows, const uint32_t *subset, int64_t n, DTransList &out, cudaStream_t st);

}  // namespace lrbk
