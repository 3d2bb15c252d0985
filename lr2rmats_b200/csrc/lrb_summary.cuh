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
};

void launch_summary_count(const SummaryArgs &a, unsigned long long *n_elems, cudaStream_t st);
// all sets + BED rows; the exon chain runs on st_exon (pass st for a single stream)
void launch_summary_sets(const SummaryArgs &a, const uint32_t *cls, int64_t n_rows, uint64_t *tile_state, uint32_t *ticket, uint64_t *bed_total,
                         cudaStream_t st, cudaStream_t st_exon, cudaEvent_t ev_fork, cudaEvent_t ev_join);

// from lrb_update.cu
void launch_rows_as_list(const DRows &rows, const uint32_t *subset, int64_t n, DTransList &out, cudaStream_t st);

}  // namespace lrbk
