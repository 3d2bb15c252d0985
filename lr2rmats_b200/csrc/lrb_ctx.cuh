// lrb_ctx.cuh -- the context behind the C ABI (private to the library): device / pinned buffers, streams, stage state and
// the small helpers every stage uses.  Shared by lrb_capi.cu (single-GPU stages) and lrb_multi.cu (NCCL layer).
#pragma once
#include <algorithm>
#include <cstdint>
#include <cstring>
#include <string>
#include "lrb_common.cuh"
#include "lrb_kernels.cuh"
#include "lrb_summary.cuh"

namespace lrbk {

extern int64_t g_launches_update, g_launches_summary, g_launches_sort;

struct Buf {                                        // device buffer, grow-only, contents not preserved on growth
    void *p = nullptr; size_t cap = 0;
    bool ensure(size_t bytes)
    {
        if (bytes <= cap) return true;
        if (p) cudaFree(p);
        size_t nc = bytes + bytes / 4 + 256;
        if (cudaMalloc(&p, nc) != cudaSuccess) { p = nullptr; cap = 0; cudaGetLastError(); return false; }
        cap = nc; return true;
    }
    void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
    template <class T> T *as() const { return (T *)p; }
};
struct PBuf {                                       // pinned host buffer
    void *p = nullptr; size_t cap = 0;
    bool ensure(size_t bytes)
    {
        if (bytes <= cap) return true;
        if (p) cudaFreeHost(p);
        size_t nc = bytes + bytes / 4 + 256;
        if (cudaMallocHost(&p, nc) != cudaSuccess) { p = nullptr; cap = 0; cudaGetLastError(); return false; }
        cap = nc; return true;
    }
    void release() { if (p) cudaFreeHost(p); p = nullptr; cap = 0; }
    template <class T> T *as() const { return (T *)p; }
};

struct MergeBufs {                                  // scratch + output of one merge fold
    Buf keys, head, locus_start, locus_cnt, dropped, rep, lstart, evmask, samemask, hard, desc, relsym, ckey, cmin, cord, clist, crow, fb_list, fb_cnt;
    Buf w_cand, w_cov, w_tid, w_start, w_end, w_fs, w_le;
    Buf o_cand, o_cov, o_tid, o_start, o_end, o_fs, o_le;
    Buf c_tid, c_start, c_end, c_rev, c_n, c_fs, c_le, c_gbeg, c_hash, c_j0, c_sig;
    int64_t n_out = 0, n_loci = 0;
    MergeArgs args;                                 // of the last fold (refolds of the split-piece rounds reuse them)
};

struct MultiState;                                  // lrb_multi.cu

}  // namespace lrbk

struct lrb_ctx {
    using Buf = lrbk::Buf; using PBuf = lrbk::PBuf; using MergeBufs = lrbk::MergeBufs; using DAnno = lrbk::DAnno; using DSj = lrbk::DSj; using DRmIndex = lrbk::DRmIndex;
    using DBatch = lrbk::DBatch; using DRows = lrbk::DRows; using DExons = lrbk::DExons; using DTransList = lrbk::DTransList;
    lrbk::MultiState *multi = nullptr;              // NCCL communicator + gather buffers (lrb_comm_init)

    int device = 0; cudaStream_t st = nullptr; std::string err;
    // side stream: work of a stage that is independent of its main chain (the class folds of the summary) runs here,
    // forked / joined with events, on its own look-back state
    cudaStream_t st2 = nullptr; cudaEvent_t ev_fork = nullptr, ev_join = nullptr; bool side_stream = true;
    cudaStream_t st3 = nullptr; cudaEvent_t ev_fork3 = nullptr, ev_join3 = nullptr; bool sum_split = true;   // exon chain of the summary sets
    // tables
    DAnno anno; DSj sj; DRmIndex rm;
    Buf a_tid, a_start, a_end, a_gene, a_rev, a_off, a_es, a_ee, a_pmax, a_mono;
    Buf s_tid, s_don, s_acc, s_u, s_m, s_pmax, s_dkey;
    Buf r_gtid, r_goff, r_start, r_pmax;
    // batch
    DBatch b; Buf b_tid, b_pos, b_lq, b_nm, b_flag, b_xs, b_qh, b_coff, b_cig;
    bool have_batch = false;
    // filter
    Buf f_pass, f_score, f_intron, f_keep_row_mask, f_keep_rec_mask, f_keep_idx, f_keep_rows;
    int64_t n_pass = 0, n_keep = 0; bool have_filter = false;
    // rows + exons
    DRows rows, rows2, rows3; DRows *cur = nullptr; DExons ex;      // rows3: the coordinate-sorted copy made by lrb_rows_sort
    Buf r_read, r_tid, r_rs, r_re, r_rev, r_beg, r_n, r_nonmono;
    Buf q_read, q_tid, q_rs, q_re, q_rev, q_beg, q_n;
    Buf s_read, s_rtid, s_rs, s_re, s_rev, s_beg, s_n, s_key0, s_key1, s_idx0, s_idx1, s_hist;   // lrb_rows_sort
    Buf e_s, e_e, e_f;
    bool have_exons = false, rows_compact = false;
    // update
    Buf u_cls, u_ref, u_nnovel, u_noff, u_mk, u_mu, u_ck, u_cr, u_cu, u_cn, u_known, u_unrecog, u_sub;
    DTransList novel; Buf n_row, n_lo, n_cnt, n_piece;
    DTransList tmp_list; Buf t_row, t_lo, t_cnt, t_piece;
    MergeBufs mg, mg2;
    int64_t n_known = 0, n_unrecog = 0, novel_cap_hint = 0; bool have_update = false, have_unique = false;
    int32_t summary[LRB_S_COUNT];
    bool force_single_fold = false, test_small_novel_cap = false, xlocus_seen = false, want_kg_pairs = false; int64_t n_xlocus_replays = 0, n_xlocus_pieces = 0, n_kg_pairs = 0; Buf kg_pairs;
    Buf xl_key, xl_min, xl_max, xl_cnt, xl_forced; lrbk::XlArgs xl;      // split pieces that meet another chromosome (xl_* kernels)
    // summary
    Buf h_khi, h_klo, h_min, h_score, y_barcnt, y_barseg, y_genebar, y_bedcnt, y_bedoff, y_counts, y_nelem;
    Buf bd_tid, bd_s, bd_e, bd_sc, bd_ty, bd_rv; int64_t n_bed = 0;
    // bam2sj
    Buf j_cnt, j_off, j_uq, j_tid, j_don, j_acc, j_u, j_head, j_hpos, jo_tid, jo_don, jo_acc, jo_u, jo_m;
    // unique
    Buf q_shared; int64_t n_shared = 0;
    // look-back state, small device scalars and their pinned mirror
    Buf tile_state, tile_state2, scalars; PBuf h_scalars;
    // pinned result buffers
    PBuf p[48];
    Buf tb_name, tb_piece, tb_ttid, tb_tstart, tb_tend, tb_trev, tb_etid, tb_erev, tb_cov, tb_ref, tb_cnt, tb_off, tb_es, tb_ee, tb_flag;
    // timing
    bool timing = false; cudaEvent_t ev[12]; cudaEvent_t marks[8]; float ms[LRB_T_COUNT]; int64_t launches0 = 0, launches_last = 0;
    lrb_update_params last_up;
};

namespace lrbk {

int build_update_table(lrb_ctx *c, int64_t *n_exon_out, bool with_flags);      // lrb_capi.cu
void multi_release(lrb_ctx *c);                                                   // lrb_multi.cu

inline int64_t total_launches() { return lrbk::count_launches() + lrbk::g_launches_update + lrbk::g_launches_summary + lrbk::g_launches_sort; }

inline int fail(lrb_ctx *c, int code, const std::string &msg) { c->err = msg; return code; }
#define CK(call)                                                                                               \
    do { cudaError_t e_ = (call); if (e_ != cudaSuccess) return fail(c, LRB_E_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_)); } while (0)
#define NEED(buf, bytes) do { if (!(buf).ensure(bytes)) return fail(c, LRB_E_NOMEM, "device allocation failed: " #buf); } while (0)
#define NEEDP(buf, bytes) do { if (!(buf).ensure(bytes)) return fail(c, LRB_E_NOMEM, "pinned allocation failed: " #buf); } while (0)

// device scalars: [0..31] uint64 totals, then uint32 ticket, err flags.  Slots 0..7 are scratch of the stage that is
// running; the update stage parks its results in fixed slots so that ONE copy brings them all to the host:
enum { T_NOVEL = 8, T_KNOWN = 9, T_UNREC = 10, T_LOCI = 11, T_UPD = 12, T_LOCI2 = 13, T_UPD2 = 14, T_NELEM = 15, T_BED = 16, T_SLOTS = 32 };
inline uint64_t *d_totals(lrb_ctx *c) { return c->scalars.as<uint64_t>(); }
inline uint32_t *d_ticket(lrb_ctx *c) { return (uint32_t *)(c->scalars.as<uint64_t>() + T_SLOTS); }
inline uint32_t *d_err(lrb_ctx *c) { return d_ticket(c) + 1; }
inline uint32_t *d_ticket2(lrb_ctx *c) { return d_ticket(c) + 2; }          // ticket of the side stream

inline int read_totals(lrb_ctx *c, uint64_t *out, int n)
{
    CK(cudaMemcpyAsync(c->h_scalars.p, d_totals(c), 8 * (size_t)n, cudaMemcpyDeviceToHost, c->st));
    CK(cudaStreamSynchronize(c->st));
    memcpy(out, c->h_scalars.p, 8 * (size_t)n);
    return LRB_OK;
}

inline int ensure_tiles(lrb_ctx *c, int64_t n_items)
{
    int64_t tiles = n_items / 8 + 1024;             // generous for every tiling used (>= n/256/8, n/reads_per_tile>=8)
    NEED(c->tile_state, (size_t)tiles * 8);
    return LRB_OK;
}

inline void tick(lrb_ctx *c, int k) { if (c->timing) cudaEventRecord(c->ev[k], c->st); }

template <class T> int h2d(lrb_ctx *c, Buf &dst, const T *src, size_t n)
{
    NEED(dst, std::max<size_t>(n, 1) * sizeof(T));
    if (n) CK(cudaMemcpyAsync(dst.p, src, n * sizeof(T), cudaMemcpyHostToDevice, c->st));
    return LRB_OK;
}
template <class T> int d2h(lrb_ctx *c, PBuf &dst, const T *src, size_t n)
{
    NEEDP(dst, std::max<size_t>(n, 1) * sizeof(T));
    if (n) CK(cudaMemcpyAsync(dst.p, src, n * sizeof(T), cudaMemcpyDeviceToHost, c->st));
    return LRB_OK;
}

}  // namespace lrbk
