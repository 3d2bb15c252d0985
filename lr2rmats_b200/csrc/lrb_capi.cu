// lrb_capi.cu -- the C ABI of liblr2rmats_b200.so (include/lr2rmats_b200.h): context, HBM residency, stage
// orchestration on one CUDA stream, pinned result buffers.  No compute happens on the host here beyond building the
// small lookup indices of the replicated tables (prefix-max keys, the remove-GTF index) at upload time.
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <string>
#include <vector>
#include "lrb_ctx.cuh"

using namespace lrbk;

namespace {

int setup_rows(lrb_ctx *c, DRows &r, Buf &read, Buf &tid, Buf &rs, Buf &re, Buf &rev, Buf &beg, Buf &cnt, int64_t cap)
{
    size_t n = (size_t)std::max<int64_t>(cap, 1);
    NEED(read, n * 4); NEED(tid, n * 4); NEED(rs, n * 4); NEED(re, n * 4); NEED(rev, n); NEED(beg, n * 4); NEED(cnt, n * 4);
    r.cap = cap; r.read_idx = read.as<uint32_t>(); r.tid = tid.as<int32_t>(); r.start = rs.as<int32_t>(); r.end = re.as<int32_t>();
    r.is_rev = rev.as<uint8_t>(); r.ex_beg = beg.as<uint32_t>(); r.ex_n = cnt.as<uint32_t>();
    return LRB_OK;
}
int setup_exons(lrb_ctx *c, int64_t cap)
{
    size_t n = (size_t)std::max<int64_t>(cap, 1);
    NEED(c->e_s, n * 4); NEED(c->e_e, n * 4); NEED(c->e_f, n);
    c->ex.cap = (int64_t)std::min(c->e_s.cap / 4, std::min(c->e_e.cap / 4, c->e_f.cap));
    c->ex.es = c->e_s.as<int32_t>(); c->ex.ee = c->e_e.as<int32_t>(); c->ex.flag = c->e_f.as<uint8_t>();
    return LRB_OK;
}

// the scan stage in one of its three modes; on return rows.n / ex.n are known on the host
// by_record (fused pipeline): rows are written at their record index and the totals stay on the device (slots 0/1) -- the
// caller reads them together with its own counts and repeats the pass if the exon pool was too small
int run_scan(lrb_ctx *c, int mode, const lrb_filter_params *fp, const lrb_exon_params *ep, const uint8_t *sel_mask, bool by_record = false)
{
    const int64_t n = c->b.n;
    int rc;
    if ((rc = ensure_tiles(c, n)) != LRB_OK) return rc;
    if ((rc = setup_rows(c, c->rows, c->r_read, c->r_tid, c->r_rs, c->r_re, c->r_rev, c->r_beg, c->r_n, n)) != LRB_OK) return rc;
    if (mode != 1) { NEED(c->f_pass, (size_t)n + 1); NEED(c->f_score, (size_t)n * 4 + 4); NEED(c->f_intron, (size_t)n * 4 + 4); }
    if (n == 0) { c->rows.n = 0; c->ex.n = 0; if (mode != 0 && (rc = setup_exons(c, 16)) != LRB_OK) return rc; return LRB_OK; }
    const double avg = (double)c->b.n_cigar / (double)n;
    // short CIGARs (Iso-Seq like): thread per read on staged tiles of 256 reads; long CIGARs (ONT like): the streaming kernel
    const bool stream_mode = avg > 48.0;
    const int R = stream_mode ? stream_reads_per_tile() : 256;
    int stage_words = (int)(R * avg * 1.3) + 256; stage_words = (stage_words + 255) & ~255; stage_words = std::max(2048, std::min(12288, stage_words));
    if (stream_mode) stage_words = 0;
    const int n_tiles = (int)((n + R - 1) / R);
    if (mode != 0) {
        int64_t bound = c->b.n_cigar + n + 16, est = c->b.n_cigar / 2 + n + 4096;
        if ((rc = setup_exons(c, std::min(bound, std::max(est, c->ex.cap)))) != LRB_OK) return rc;
    }
    for (int attempt = 0; attempt < 2; ++attempt) {
        ScanArgs a{};
        a.b = c->b; if (fp) a.fp = *fp; if (ep) a.ep = *ep; a.rm = c->rm; a.mode = mode; a.sel_mask = sel_mask;
        a.pass = c->f_pass.as<uint8_t>(); a.score = c->f_score.as<int32_t>(); a.intron_n = c->f_intron.as<int32_t>();
        a.rows = c->rows; a.ex = c->ex; a.tile_state = c->tile_state.as<uint64_t>(); a.ticket = d_ticket(c); a.totals = d_totals(c);
        a.reads_per_tile = R; a.stage_words = stage_words; a.rows_by_record = by_record ? 1 : 0;
        CK(cudaMemsetAsync(c->tile_state.p, 0, (size_t)n_tiles * 8, c->st));
        CK(cudaMemsetAsync(c->scalars.p, 0, 8 * 8, c->st)); CK(cudaMemsetAsync(d_ticket(c), 0, 4, c->st));
        size_t smem = (size_t)stage_words * 4;
        tick(c, 8);
        launch_cigar_scan(a, n_tiles, stream_mode, smem, c->st);
        tick(c, 9);
        CK(cudaGetLastError());
        if (by_record) return LRB_OK;
        uint64_t t[2];
        if ((rc = read_totals(c, t, 2)) != LRB_OK) return rc;
        c->rows.n = (int64_t)t[0]; c->ex.n = (int64_t)t[1];
        if (c->timing) cudaEventElapsedTime(&c->ms[LRB_T_K_SCAN], c->ev[8], c->ev[9]);
        if (mode == 0 || c->ex.n <= c->ex.cap) break;
        if (attempt == 1) return fail(c, LRB_E_NOMEM, "exon pool still too small after regrow");
        if ((rc = setup_exons(c, c->ex.n + 16)) != LRB_OK) return rc;          // exact size known now: walk again
    }
    return LRB_OK;
}

int run_select(lrb_ctx *c, const lrb_filter_params *fp)
{
    const int64_t n = c->b.n, np = c->rows.n;
    NEED(c->f_keep_row_mask, (size_t)std::max<int64_t>(np, 1)); NEED(c->f_keep_rec_mask, (size_t)std::max<int64_t>(n, 1));
    NEED(c->f_keep_idx, (size_t)std::max<int64_t>(np, 1) * 4); NEED(c->f_keep_rows, (size_t)std::max<int64_t>(np, 1) * 4);
    CK(cudaMemsetAsync(c->f_keep_row_mask.p, 0, (size_t)std::max<int64_t>(np, 1), c->st));
    CK(cudaMemsetAsync(c->f_keep_rec_mask.p, 0, (size_t)std::max<int64_t>(n, 1), c->st));
    launch_select_runs(c->b, c->rows.read_idx, np, c->f_score.as<int32_t>(), c->f_intron.as<int32_t>(), *fp,
                       c->f_keep_row_mask.as<uint8_t>(), c->f_keep_rec_mask.as<uint8_t>(), c->st);
    launch_compact_mask(c->f_keep_row_mask.as<uint8_t>(), np, c->rows.read_idx, c->f_keep_idx.as<uint32_t>(), c->f_keep_rows.as<uint32_t>(),
                        c->tile_state.as<uint64_t>(), d_ticket(c), d_totals(c), c->st);
    CK(cudaGetLastError());
    uint64_t t; int rc = read_totals(c, &t, 1); if (rc) return rc;
    c->n_pass = np; c->n_keep = (int64_t)t;
    return LRB_OK;
}

int setup_merge(lrb_ctx *c, MergeBufs &m, int64_t n_cand)
{
    size_t n = (size_t)std::max<int64_t>(n_cand, 1);
    NEED(m.keys, n * 8); NEED(m.head, n); NEED(m.locus_start, (n + 1) * 4); NEED(m.locus_cnt, n * 4); NEED(m.dropped, n);
    NEED(m.rep, n * 4); NEED(m.lstart, n * 4); NEED(m.evmask, n * 8); NEED(m.samemask, n * 8); NEED(m.hard, n); NEED(m.desc, n * 2); NEED(m.relsym, n * 8); NEED(m.ckey, (2 * n + 64) * 8); NEED(m.cmin, (2 * n + 64) * 4); NEED(m.cord, n); NEED(m.clist, n * 4); NEED(m.crow, n * 16); NEED(m.fb_list, n * 8); NEED(m.fb_cnt, 64);
    Buf *w[] = {&m.w_cand, &m.w_cov, &m.w_tid, &m.w_start, &m.w_end, &m.w_fs, &m.w_le, &m.o_cand, &m.o_cov, &m.o_tid, &m.o_start, &m.o_end, &m.o_fs, &m.o_le};
    for (Buf *b : w) NEED(*b, n * 4);
    Buf *cb[] = {&m.c_tid, &m.c_start, &m.c_end, &m.c_rev, &m.c_n, &m.c_fs, &m.c_le, &m.c_gbeg};
    for (Buf *b : cb) NEED(*b, n * 4);
    NEED(m.c_hash, n * 8); NEED(m.c_j0, n * 8); NEED(m.c_sig, n * 8);
    return LRB_OK;
}
DMerged merged_view(Buf &cand, Buf &cov, Buf &tid, Buf &st, Buf &en, Buf &fs, Buf &le, int64_t n)
{
    DMerged d; d.n = n; d.cap = n; d.cand = cand.as<uint32_t>(); d.cov = cov.as<int32_t>(); d.tid = tid.as<int32_t>(); d.start = st.as<int32_t>();
    d.end = en.as<int32_t>(); d.fs = fs.as<int32_t>(); d.le = le.as<int32_t>();
    return d;
}

// merge fold over `list`, asynchronous.  n_cand is the number of candidates, or (n_cand_dev != nullptr) the host's upper
// bound of a count that lives on the device.  kls != nullptr: the list carries four independent sub-streams (class id per
// candidate) folded in one pass and only the survivors per sub-stream are counted (class_alive); else the survivors are
// compacted into m.o_*.  totals[0] <- number of loci, totals[1] <- number of survivors (device).
int run_merge_async(lrb_ctx *c, MergeBufs &m, const DTransList &list, int64_t n_cand, const uint64_t *n_cand_dev, const lrb_update_params &up,
                    uint64_t *totals, const uint8_t *kls = nullptr, uint32_t *class_alive = nullptr, bool time_fold = false, bool side = false,
                    bool single_locus = false, bool xl = false)
{
    int rc;
    cudaStream_t st = side ? c->st2 : c->st;
    if ((rc = setup_merge(c, m, n_cand)) != LRB_OK) return rc;
    if (side) { NEED(c->tile_state2, (size_t)(n_cand / 8 + 1024) * 8); }
    else if ((rc = ensure_tiles(c, n_cand)) != LRB_OK) return rc;
    m.n_out = 0; m.n_loci = 0;
    CK(cudaMemsetAsync(totals, 0, 16, st));
    if (n_cand == 0) return LRB_OK;
    MergeArgs a{};
    a.rows = *c->cur; a.ex = c->ex; a.up = up; a.list = list; a.n_cand = n_cand; a.n_cand_dev = n_cand_dev;
    a.keys = m.keys.as<uint64_t>(); a.head = m.head.as<uint8_t>(); a.locus_start = m.locus_start.as<uint32_t>(); a.locus_cnt = m.locus_cnt.as<uint32_t>();
    a.dropped = m.dropped.as<uint8_t>();
    a.rep = m.rep.as<uint32_t>(); a.lstart = m.lstart.as<uint32_t>(); a.evmask = m.evmask.as<uint64_t>(); a.samemask = m.samemask.as<uint64_t>(); a.hard = m.hard.as<uint8_t>(); a.desc = m.desc.as<uint16_t>(); a.relsym = m.relsym.as<uint64_t>(); a.ckey = m.ckey.as<uint64_t>(); a.cmin = m.cmin.as<uint32_t>(); a.cord = m.cord.as<uint8_t>(); a.clist = m.clist.as<uint32_t>(); a.crow = m.crow.as<uint64_t>(); a.fb_list = m.fb_list.as<uint32_t>(); a.fb_cnt = m.fb_cnt.as<uint32_t>();
    a.work = merged_view(m.w_cand, m.w_cov, m.w_tid, m.w_start, m.w_end, m.w_fs, m.w_le, n_cand);
    a.out = merged_view(m.o_cand, m.o_cov, m.o_tid, m.o_start, m.o_end, m.o_fs, m.o_le, n_cand);
    a.cd.tid = m.c_tid.as<int32_t>(); a.cd.start = m.c_start.as<int32_t>(); a.cd.end = m.c_end.as<int32_t>(); a.cd.rev = m.c_rev.as<int32_t>();
    a.cd.n = m.c_n.as<int32_t>(); a.cd.fs = m.c_fs.as<int32_t>(); a.cd.le = m.c_le.as<int32_t>(); a.cd.gbeg = m.c_gbeg.as<uint32_t>();
    a.cd.hash = m.c_hash.as<uint64_t>(); a.cd.j0 = m.c_j0.as<uint64_t>(); a.cd.sig = m.c_sig.as<uint64_t>();
    a.tile_state = side ? c->tile_state2.as<uint64_t>() : c->tile_state.as<uint64_t>(); a.ticket = side ? d_ticket2(c) : d_ticket(c); a.totals = totals;
    a.kls = kls; a.class_alive = class_alive; a.single_locus = single_locus ? 1 : 0;
    if (xl) {
        // split pieces that meet another chromosome (xl_* kernels): marks per candidate, a key set sized for the pieces' junctions
        // (pieces are a small share of the candidates), lists in buffers the fold is done with by then
        const uint64_t tcap = std::max<uint64_t>(1u << 16, (uint64_t)n_cand / 2);
        NEED(c->xl_key, tcap * 8); NEED(c->xl_min, tcap * 4); NEED(c->xl_max, tcap * 4); NEED(c->xl_cnt, 64);
        a.forced = c->xl_forced.as<uint8_t>();
        c->xl = XlArgs{c->xl_key.as<unsigned long long>(), c->xl_min.as<uint32_t>(), c->xl_max.as<uint32_t>(), tcap, a.clist, a.lstart, (uint32_t)std::min<int64_t>(n_cand, 1 << 20),
                       a.relsym, c->xl_cnt.as<uint32_t>()};
    }
    m.args = a;
    launch_merge_prepare(a, st);
    if (time_fold) tick(c, 10);
    launch_merge_fold(a, st);                        // locus count is consumed on the device: no host round trip
    if (time_fold) tick(c, 11);
    if (xl) launch_xlocus_detect(a, c->xl, st);
    if (kls) launch_merge_class_counts(a, st);
    else launch_merge_finish(a, st);
    CK(cudaGetLastError());
    return LRB_OK;
}

int run_merge(lrb_ctx *c, MergeBufs &m, const DTransList &list, int64_t n_cand, const lrb_update_params &up, bool single_locus = false)
{
    int rc;
    if ((rc = run_merge_async(c, m, list, n_cand, nullptr, up, d_totals(c), nullptr, nullptr, true, false, single_locus)) != LRB_OK) return rc;
    uint64_t t[2];
    if ((rc = read_totals(c, t, 2)) != LRB_OK) return rc;
    m.n_loci = (int64_t)t[0]; m.n_out = (int64_t)t[1];
    if (c->timing && n_cand) cudaEventElapsedTime(&c->ms[LRB_T_K_FOLD], c->ev[10], c->ev[11]);
    return LRB_OK;
}

int setup_list(lrb_ctx *c, DTransList &l, Buf &row, Buf &lo, Buf &cnt, Buf &piece, int64_t n)
{
    size_t k = (size_t)std::max<int64_t>(n, 1);
    NEED(row, k * 4); NEED(lo, k * 4); NEED(cnt, k * 4); NEED(piece, k * 4);
    l.n = n; l.cap = n; l.row = row.as<uint32_t>(); l.lo = lo.as<uint32_t>(); l.cnt = cnt.as<uint32_t>(); l.piece = piece.as<int32_t>();
    return LRB_OK;
}

}  // namespace

// ====================================================================================================== C ABI
extern "C" {

const char *lrb_version(void) { return "lr2rmats_b200 0.1 (sm_100a)"; }

int lrb_ctx_create(int device, lrb_ctx **out)
{
    if (!out) return LRB_E_ARG;
    *out = nullptr;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0) { cudaGetLastError(); return LRB_E_NODEVICE; }
    if (device < 0 || device >= ndev) return LRB_E_ARG;
    if (cudaSetDevice(device) != cudaSuccess) { cudaGetLastError(); return LRB_E_CUDA; }
    lrb_ctx *c = new lrb_ctx();
    c->device = device;
    if (cudaStreamCreateWithFlags(&c->st, cudaStreamNonBlocking) != cudaSuccess) { delete c; return LRB_E_CUDA; }
    if (cudaStreamCreateWithFlags(&c->st2, cudaStreamNonBlocking) != cudaSuccess) { cudaStreamDestroy(c->st); delete c; return LRB_E_CUDA; }
    cudaEventCreateWithFlags(&c->ev_fork, cudaEventDisableTiming); cudaEventCreateWithFlags(&c->ev_join, cudaEventDisableTiming);
    if (cudaStreamCreateWithFlags(&c->st3, cudaStreamNonBlocking) != cudaSuccess) { cudaStreamDestroy(c->st2); cudaStreamDestroy(c->st); delete c; return LRB_E_CUDA; }
    cudaEventCreateWithFlags(&c->ev_fork3, cudaEventDisableTiming); cudaEventCreateWithFlags(&c->ev_join3, cudaEventDisableTiming);
    { const char *e = getenv("LRB_SUM_SPLIT"); if (e) c->sum_split = atoi(e) != 0; }
    { const char *e = getenv("LRB_SIDE_STREAM"); if (e) c->side_stream = atoi(e) != 0; }
    { const char *e = getenv("LRB_TEST_SMALL_NOVEL_CAP"); if (e) c->test_small_novel_cap = atoi(e) != 0; }      // test hook: novel_T sized too small on the first attempt
    { const char *e = getenv("LRB_FORCE_SINGLE_FOLD"); if (e) c->force_single_fold = atoi(e) != 0; }    // test hook: updated_T always folded by the one-locus replay
    if (!c->scalars.ensure(512) || !c->h_scalars.ensure(512)) { delete c; return LRB_E_NOMEM; }
    cudaMemsetAsync(c->scalars.p, 0, 512, c->st);
    for (int i = 0; i < 12; ++i) cudaEventCreate(&c->ev[i]);
    for (int i = 0; i < 8; ++i) cudaEventCreate(&c->marks[i]);
    memset(c->ms, 0, sizeof c->ms); memset(c->summary, 0, sizeof c->summary);
    c->launches0 = total_launches();
    *out = c;
    return LRB_OK;
}

void lrb_ctx_destroy(lrb_ctx *c)
{
    if (!c) return;
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->st); cudaStreamSynchronize(c->st2); cudaStreamSynchronize(c->st3);
    Buf *bufs[] = {&c->a_tid, &c->a_start, &c->a_end, &c->a_gene, &c->a_rev, &c->a_off, &c->a_es, &c->a_ee, &c->a_pmax, &c->a_mono, &c->s_tid, &c->s_don, &c->s_acc,
                   &c->s_u, &c->s_m, &c->s_pmax, &c->s_dkey, &c->r_gtid, &c->r_goff, &c->r_start, &c->r_pmax, &c->b_tid, &c->b_pos, &c->b_lq, &c->b_nm,
                   &c->b_flag, &c->b_xs, &c->b_qh, &c->b_coff, &c->b_cig, &c->f_pass, &c->f_score, &c->f_intron, &c->f_keep_row_mask, &c->f_keep_rec_mask,
                   &c->f_keep_idx, &c->f_keep_rows, &c->r_read, &c->r_tid, &c->r_rs, &c->r_re, &c->r_rev, &c->r_beg, &c->r_n, &c->r_nonmono, &c->q_read, &c->q_tid,
                   &c->q_rs, &c->q_re, &c->q_rev, &c->q_beg, &c->q_n, &c->e_s, &c->e_e, &c->e_f, &c->u_cls, &c->u_ref, &c->u_nnovel, &c->u_noff, &c->u_mk,
                   &c->u_mu, &c->u_ck, &c->u_cr, &c->u_cu, &c->u_cn, &c->u_known, &c->u_unrecog, &c->u_sub, &c->n_row, &c->n_lo, &c->n_cnt, &c->n_piece,
                   &c->t_row, &c->t_lo, &c->t_cnt, &c->t_piece, &c->h_khi, &c->h_klo, &c->h_min, &c->h_score, &c->y_barcnt, &c->y_barseg, &c->y_genebar,
                   &c->y_bedcnt, &c->y_bedoff, &c->y_counts, &c->y_nelem, &c->bd_tid, &c->bd_s, &c->bd_e, &c->bd_sc, &c->bd_ty, &c->bd_rv, &c->q_shared, &c->tb_name, &c->tb_piece, &c->tb_ttid, &c->tb_tstart, &c->tb_tend, &c->tb_trev, &c->tb_etid, &c->tb_erev, &c->tb_cov, &c->tb_ref, &c->tb_cnt, &c->tb_off, &c->tb_es, &c->tb_ee, &c->tb_flag,
                   &c->s_read, &c->s_rtid, &c->s_rs, &c->s_re, &c->s_rev, &c->s_beg, &c->s_n, &c->s_key0, &c->s_key1, &c->s_idx0, &c->s_idx1, &c->s_hist,
                   &c->tile_state, &c->tile_state2, &c->scalars, &c->kg_pairs, &c->xl_key, &c->xl_min, &c->xl_max, &c->xl_cnt, &c->xl_forced,
                   &c->j_cnt, &c->j_off, &c->j_uq, &c->j_tid, &c->j_don, &c->j_acc, &c->j_u, &c->j_head, &c->j_hpos, &c->jo_tid, &c->jo_don, &c->jo_acc, &c->jo_u, &c->jo_m};
    for (Buf *b : bufs) b->release();
    for (MergeBufs *m : {&c->mg, &c->mg2}) {
        Buf *w[] = {&m->keys, &m->head, &m->locus_start, &m->locus_cnt, &m->dropped, &m->rep, &m->lstart, &m->evmask, &m->samemask, &m->hard, &m->desc, &m->relsym, &m->ckey, &m->cmin, &m->cord, &m->clist, &m->crow, &m->fb_list, &m->fb_cnt, &m->w_cand, &m->w_cov, &m->w_tid, &m->w_start, &m->w_end, &m->w_fs,
                    &m->w_le, &m->o_cand, &m->o_cov, &m->o_tid, &m->o_start, &m->o_end, &m->o_fs, &m->o_le,
                    &m->c_tid, &m->c_start, &m->c_end, &m->c_rev, &m->c_n, &m->c_fs, &m->c_le, &m->c_gbeg, &m->c_hash, &m->c_j0, &m->c_sig};
        for (Buf *b : w) b->release();
    }
    lrbk::multi_release(c);
    for (PBuf &p : c->p) p.release();
    c->h_scalars.release();
    for (int i = 0; i < 12; ++i) cudaEventDestroy(c->ev[i]);
    for (int i = 0; i < 8; ++i) cudaEventDestroy(c->marks[i]);
    cudaEventDestroy(c->ev_fork); cudaEventDestroy(c->ev_join);
    cudaEventDestroy(c->ev_fork3); cudaEventDestroy(c->ev_join3); cudaStreamDestroy(c->st3);
    cudaStreamDestroy(c->st2); cudaStreamDestroy(c->st);
    delete c;
}

const char *lrb_last_error(const lrb_ctx *c) { return c ? c->err.c_str() : "null context"; }

// -------------------------------------------------------------------------------------------------- tables
int lrb_anno_upload(lrb_ctx *c, const lrb_anno *a)
{
    if (!c) return LRB_E_ARG;
    CK(cudaSetDevice(c->device));
    c->anno = DAnno{};
    if (!a || a->n_trans == 0) return LRB_OK;
    const int32_t n = a->n_trans;
    std::vector<uint64_t> pm((size_t)n);
    uint64_t run = 0;
    for (int32_t i = 0; i < n; ++i) {               // P_j, SURVEY App. B.1
        uint64_t k = ((uint64_t)(uint32_t)(a->tid[i] + 1) << 32) | (uint32_t)a->end[i];
        run = std::max(run, k); pm[(size_t)i] = run;
    }
    std::vector<uint8_t> mono((size_t)n);
    for (int32_t i = 0; i < n; ++i) {
        uint8_t m = 1;
        for (uint32_t k = a->exon_off[i] + 1; k < a->exon_off[i + 1]; ++k)
            if (a->exon_start[k] <= a->exon_start[k - 1] || a->exon_end[k] <= a->exon_end[k - 1]) { m = 0; break; }
        mono[(size_t)i] = m;
    }
    int rc;
    if ((rc = h2d(c, c->a_mono, mono.data(), (size_t)n))) return rc;
    if ((rc = h2d(c, c->a_tid, a->tid, (size_t)n))) return rc;
    if ((rc = h2d(c, c->a_start, a->start, (size_t)n))) return rc;
    if ((rc = h2d(c, c->a_end, a->end, (size_t)n))) return rc;
    if ((rc = h2d(c, c->a_gene, a->gene, (size_t)n))) return rc;
    if ((rc = h2d(c, c->a_rev, a->is_rev, (size_t)n))) return rc;
    if ((rc = h2d(c, c->a_off, a->exon_off, (size_t)n + 1))) return rc;
    if ((rc = h2d(c, c->a_es, a->exon_start, (size_t)a->n_exon))) return rc;
    if ((rc = h2d(c, c->a_ee, a->exon_end, (size_t)a->n_exon))) return rc;
    if ((rc = h2d(c, c->a_pmax, pm.data(), (size_t)n))) return rc;
    CK(cudaStreamSynchronize(c->st));               // pm is a local
    c->anno.n = n; c->anno.n_exon = a->n_exon; c->anno.tid = c->a_tid.as<int32_t>(); c->anno.start = c->a_start.as<int32_t>();
    c->anno.end = c->a_end.as<int32_t>(); c->anno.gene = c->a_gene.as<int32_t>(); c->anno.is_rev = c->a_rev.as<uint8_t>();
    c->anno.exon_off = c->a_off.as<uint32_t>(); c->anno.es = c->a_es.as<int32_t>(); c->anno.ee = c->a_ee.as<int32_t>();
    c->anno.pmax_key = c->a_pmax.as<uint64_t>(); c->anno.mono = c->a_mono.as<uint8_t>();
    return LRB_OK;
}

int lrb_rm_upload(lrb_ctx *c, const lrb_anno *rmt)
{
    if (!c) return LRB_E_ARG;
    CK(cudaSetDevice(c->device));
    c->rm = DRmIndex{};
    if (!rmt || rmt->n_trans == 0) return LRB_OK;
    // entries visible to remove_overlap()'s early exit (bam_filter.c:54-57): for a read on tid t the scan stops after the
    // first entry with tid > t, so only entries in front of it exist for t.  Group them per tid, sort by start, running max end.
    const int32_t n = rmt->n_trans;
    std::map<int32_t, std::vector<std::pair<int32_t, int32_t>>> groups;
    std::vector<int32_t> pmax_tid((size_t)n);
    int32_t run = INT32_MIN;
    for (int32_t i = 0; i < n; ++i) { run = std::max(run, rmt->tid[i]); pmax_tid[(size_t)i] = run; }
    for (int32_t i = 0; i < n; ++i) {
        int32_t t = rmt->tid[i];
        // visible iff no earlier-or-equal position has prefix max tid > t  <=>  pmax_tid[i] <= t (i itself has tid t)
        if (pmax_tid[(size_t)i] > t) continue;
        groups[t].push_back({rmt->start[i], rmt->end[i]});
    }
    std::vector<int32_t> gt, go{0}, st, pe;
    for (auto &g : groups) {
        std::sort(g.second.begin(), g.second.end());
        int32_t m = INT32_MIN;
        for (auto &iv : g.second) { m = std::max(m, iv.second); st.push_back(iv.first); pe.push_back(m); }
        gt.push_back(g.first); go.push_back((int32_t)st.size());
    }
    int rc;
    if ((rc = h2d(c, c->r_gtid, gt.data(), gt.size()))) return rc;
    if ((rc = h2d(c, c->r_goff, go.data(), go.size()))) return rc;
    if ((rc = h2d(c, c->r_start, st.data(), st.size()))) return rc;
    if ((rc = h2d(c, c->r_pmax, pe.data(), pe.size()))) return rc;
    CK(cudaStreamSynchronize(c->st));
    c->rm.n_groups = (int32_t)gt.size(); c->rm.n = (int32_t)st.size(); c->rm.g_tid = c->r_gtid.as<int32_t>(); c->rm.g_off = c->r_goff.as<int32_t>();
    c->rm.start = c->r_start.as<int32_t>(); c->rm.pmax_end = c->r_pmax.as<int32_t>();
    return LRB_OK;
}

int lrb_sj_upload(lrb_ctx *c, const lrb_sj *s)
{
    if (!c) return LRB_E_ARG;
    CK(cudaSetDevice(c->device));
    c->sj = DSj{};
    if (!s || s->n == 0) return LRB_OK;
    const int64_t n = s->n;
    std::vector<uint64_t> pm((size_t)n), dk((size_t)n);
    uint64_t run = 0;
    for (int64_t i = 0; i < n; ++i) {               // Q_i, SURVEY App. B.2
        uint64_t k = ((uint64_t)(uint32_t)(s->tid[i] + 1) << 32) | (uint32_t)s->acc[i];
        run = std::max(run, k); pm[(size_t)i] = run;
        dk[(size_t)i] = ((uint64_t)(uint32_t)(s->tid[i] + 1) << 32) | (uint32_t)s->don[i];
        if (i && dk[(size_t)i] < dk[(size_t)i - 1]) return fail(c, LRB_E_ARG, "SJ table is not sorted by (tid,don,acc)");
    }
    int rc;
    if ((rc = h2d(c, c->s_tid, s->tid, (size_t)n))) return rc;
    if ((rc = h2d(c, c->s_don, s->don, (size_t)n))) return rc;
    if ((rc = h2d(c, c->s_acc, s->acc, (size_t)n))) return rc;
    if ((rc = h2d(c, c->s_u, s->uniq_c, (size_t)n))) return rc;
    if ((rc = h2d(c, c->s_m, s->multi_c, (size_t)n))) return rc;
    if ((rc = h2d(c, c->s_pmax, pm.data(), (size_t)n))) return rc;
    if ((rc = h2d(c, c->s_dkey, dk.data(), (size_t)n))) return rc;
    CK(cudaStreamSynchronize(c->st));
    c->sj.n = n; c->sj.tid = c->s_tid.as<int32_t>(); c->sj.don = c->s_don.as<int32_t>(); c->sj.acc = c->s_acc.as<int32_t>();
    c->sj.cnt_u = c->s_u.as<int32_t>(); c->sj.cnt_m = c->s_m.as<int32_t>(); c->sj.pmax_key = c->s_pmax.as<uint64_t>(); c->sj.don_key = c->s_dkey.as<uint64_t>();
    return LRB_OK;
}

// --------------------------------------------------------------------------------------------------- batch
int lrb_batch_upload(lrb_ctx *c, const lrb_batch *b)
{
    if (!c || !b || b->n < 0) return LRB_E_ARG;
    CK(cudaSetDevice(c->device));
    const size_t n = (size_t)b->n;
    const size_t nc = n ? b->cigar_off[n] : 0;
    if (b->n >= (1ll << 31) - 1) return fail(c, LRB_E_ARG, "batch too large: at most 2^31-2 records per batch");
    if (n && b->cigar_off[0] != 0) return fail(c, LRB_E_ARG, "cigar_off[0] must be 0");
    if (nc >= (1ull << 40)) return fail(c, LRB_E_ARG, "batch too large: at most 2^40 CIGAR ops per batch");
    int rc;
    if ((rc = h2d(c, c->b_tid, b->tid, n))) return rc;
    if ((rc = h2d(c, c->b_pos, b->pos, n))) return rc;
    if ((rc = h2d(c, c->b_flag, b->flag, n))) return rc;
    if ((rc = h2d(c, c->b_lq, b->l_qseq, n))) return rc;
    if ((rc = h2d(c, c->b_nm, b->nm, n))) return rc;
    if ((rc = h2d(c, c->b_xs, b->xs, n))) return rc;
    if ((rc = h2d(c, c->b_qh, b->qname_hash, n))) return rc;
    if (n) { if ((rc = h2d(c, c->b_coff, b->cigar_off, n + 1))) return rc; }
    else { NEED(c->b_coff, 8); CK(cudaMemsetAsync(c->b_coff.p, 0, 8, c->st)); }
    if ((rc = h2d(c, c->b_cig, b->cigar, nc))) return rc;
    c->b.n = b->n; c->b.n_cigar = (int64_t)nc;
    c->b.tid = c->b_tid.as<int32_t>(); c->b.pos = c->b_pos.as<int32_t>(); c->b.flag = c->b_flag.as<uint16_t>(); c->b.l_qseq = c->b_lq.as<int32_t>();
    c->b.nm = c->b_nm.as<int32_t>(); c->b.xs = c->b_xs.as<int8_t>(); c->b.qhash = c->b_qh.as<uint64_t>(); c->b.cigar_off = c->b_coff.as<uint64_t>();
    c->b.cigar = c->b_cig.as<uint32_t>();
    c->have_batch = true; c->have_filter = c->have_exons = c->have_update = c->have_unique = false;
    return LRB_OK;
}

// rows straight from exon chains (-m g input): start/end/ex_beg/ex_n derived on device
__global__ void chains_rows_kernel(DRows rows, const uint32_t *__restrict__ off, const int32_t *__restrict__ es, const int32_t *__restrict__ ee, int64_t n,
                                   uint8_t *nonmono)
{
    int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n) return;
    uint32_t lo = off[r], hi = off[r + 1];
    uint8_t nm = 0;                                  // GTF chains may be anything; CIGAR chains always ascend
    for (uint32_t k = lo + 1; k < hi; ++k) if (es[k] < es[k - 1] || ee[k] < ee[k - 1]) nm = 1;
    nonmono[r] = nm;
    rows.read_idx[r] = (uint32_t)r; rows.ex_beg[r] = lo; rows.ex_n[r] = hi - lo;
    rows.start[r] = hi > lo ? es[lo] : 0; rows.end[r] = hi > lo ? ee[hi - 1] : 0;
}

int lrb_chains_upload(lrb_ctx *c, const lrb_chains *ch)
{
    if (!c || !ch || ch->n < 0) return LRB_E_ARG;
    CK(cudaSetDevice(c->device));
    const int64_t n = ch->n; const int64_t ne = n ? ch->exon_off[n] : 0;
    int rc;
    if ((rc = setup_rows(c, c->rows, c->r_read, c->r_tid, c->r_rs, c->r_re, c->r_rev, c->r_beg, c->r_n, n))) return rc;
    if ((rc = setup_exons(c, ne + 16))) return rc;
    if ((rc = h2d(c, c->r_tid, ch->tid, (size_t)n))) return rc;
    if ((rc = h2d(c, c->r_rev, ch->is_rev, (size_t)n))) return rc;
    if ((rc = h2d(c, c->e_s, ch->exon_start, (size_t)ne))) return rc;
    if ((rc = h2d(c, c->e_e, ch->exon_end, (size_t)ne))) return rc;
    if ((rc = h2d(c, c->u_noff, ch->exon_off, (size_t)n + 1))) return rc;          // scratch for the offsets
    NEED(c->r_nonmono, (size_t)std::max<int64_t>(n, 1));
    if (n) { chains_rows_kernel<<<(unsigned)((n + 255) / 256), 256, 0, c->st>>>(c->rows, c->u_noff.as<uint32_t>(), c->ex.es, c->ex.ee, n, c->r_nonmono.as<uint8_t>()); CK(cudaGetLastError()); }
    c->rows.n = n; c->ex.n = ne; c->cur = &c->rows; c->rows_compact = true;
    c->b.n = 0; c->have_batch = false; c->have_filter = false; c->have_exons = true; c->have_update = c->have_unique = false;
    return LRB_OK;
}

// -------------------------------------------------------------------------------------------------- stages
int lrb_filter_run(lrb_ctx *c, const lrb_filter_params *p)
{
    if (!c || !p) return LRB_E_ARG;
    if (!c->have_batch) return fail(c, LRB_E_ARG, "lrb_filter_run: no batch uploaded");
    CK(cudaSetDevice(c->device));
    int64_t l0 = total_launches(); tick(c, 0);
    int rc = run_scan(c, 0, p, nullptr, nullptr); if (rc) return rc;
    if ((rc = run_select(c, p))) return rc;
    tick(c, 1);
    c->have_filter = true; c->have_exons = false; c->launches_last = total_launches() - l0;
    if (c->timing) { CK(cudaStreamSynchronize(c->st)); cudaEventElapsedTime(&c->ms[LRB_T_FILTER], c->ev[0], c->ev[1]); }
    return LRB_OK;
}

int lrb_exon_run(lrb_ctx *c, const lrb_exon_params *p, int use_keep_list)
{
    if (!c || !p) return LRB_E_ARG;
    if (!c->have_batch) return fail(c, LRB_E_ARG, "lrb_exon_run: no batch uploaded");
    if (use_keep_list && !c->have_filter) return fail(c, LRB_E_ARG, "lrb_exon_run: keep list requested but lrb_filter_run has not run");
    CK(cudaSetDevice(c->device));
    int64_t l0 = total_launches(); tick(c, 0);
    int rc = run_scan(c, 1, nullptr, p, use_keep_list ? c->f_keep_rec_mask.as<uint8_t>() : nullptr); if (rc) return rc;
    tick(c, 1);
    c->cur = &c->rows; c->rows_compact = true; c->have_exons = true; c->have_update = c->have_unique = false;
    c->launches_last = total_launches() - l0;
    if (c->timing) { CK(cudaStreamSynchronize(c->st)); cudaEventElapsedTime(&c->ms[LRB_T_EXON], c->ev[0], c->ev[1]); }
    return LRB_OK;
}

int lrb_pipeline_run(lrb_ctx *c, const lrb_filter_params *fp, const lrb_exon_params *ep)
{
    if (!c || !fp || !ep) return LRB_E_ARG;
    if (!c->have_batch) return fail(c, LRB_E_ARG, "lrb_pipeline_run: no batch uploaded");
    CK(cudaSetDevice(c->device));
    int64_t l0 = total_launches(); tick(c, 0);
    const int64_t n = c->b.n;
    int rc;
    NEED(c->f_keep_rec_mask, (size_t)std::max<int64_t>(n, 1)); NEED(c->f_keep_idx, (size_t)std::max<int64_t>(n, 1) * 4);
    if ((rc = setup_rows(c, c->rows2, c->q_read, c->q_tid, c->q_rs, c->q_re, c->q_rev, c->q_beg, c->q_n, n))) return rc;
    c->rows.n = 0; c->ex.n = 0; c->n_pass = c->n_keep = 0;
    for (int attempt = 0; n > 0; ++attempt) {
        // scan (rows at their record index) -> run selection on the record stream -> compaction + row gather: one
        // round trip for (passing, exons, kept)
        if ((rc = run_scan(c, 2, fp, ep, nullptr, true))) return rc;
        CK(cudaMemsetAsync(c->f_keep_rec_mask.p, 0, (size_t)n, c->st));
        launch_select_records(c->b, c->f_pass.as<uint8_t>(), c->f_score.as<int32_t>(), c->f_intron.as<int32_t>(), *fp, c->f_keep_rec_mask.as<uint8_t>(), c->st);
        launch_compact_gather(c->f_keep_rec_mask.as<uint8_t>(), n, c->rows, c->rows2, c->f_keep_idx.as<uint32_t>(), c->tile_state.as<uint64_t>(),
                              d_ticket(c), d_totals(c) + 2, c->st);
        CK(cudaGetLastError());
        uint64_t t[3];
        if ((rc = read_totals(c, t, 3))) return rc;
        c->n_pass = (int64_t)t[0]; c->ex.n = (int64_t)t[1]; c->n_keep = (int64_t)t[2];
        if (c->timing) cudaEventElapsedTime(&c->ms[LRB_T_K_SCAN], c->ev[8], c->ev[9]);
        if (c->ex.n <= c->ex.cap) break;
        if (attempt == 1) return fail(c, LRB_E_NOMEM, "exon pool still too small after regrow");
        if ((rc = setup_exons(c, c->ex.n + 16))) return rc;     // exact size known now: walk again
    }
    if (n == 0 && (rc = setup_exons(c, 16))) return rc;
    c->rows.n = n;                                               // record-indexed: only the passing records' rows are valid
    c->rows2.n = c->n_keep;
    tick(c, 1);
    c->cur = &c->rows2; c->rows_compact = false; c->have_filter = true; c->have_exons = true; c->have_update = c->have_unique = false;
    c->launches_last = total_launches() - l0;
    if (c->timing) { CK(cudaStreamSynchronize(c->st)); cudaEventElapsedTime(&c->ms[LRB_T_FILTER], c->ev[0], c->ev[1]); c->ms[LRB_T_EXON] = 0; }
    return LRB_OK;
}

static int check_err_flags(lrb_ctx *c, uint32_t *flags = nullptr)
{
    uint32_t e = 0;
    CK(cudaMemcpyAsync(c->h_scalars.p, d_err(c), 4, cudaMemcpyDeviceToHost, c->st));
    CK(cudaStreamSynchronize(c->st));
    memcpy(&e, c->h_scalars.p, 4);
    if (flags) *flags = e;
    if (e & 2u) return fail(c, LRB_E_UNMAPPED, "unmapped record / empty exon chain in update/unique input (the reference aborts here, bam2gtf.c:95-100)");
    if (e & 1u) return fail(c, LRB_E_UNSORTED, "reads are not sorted by (tid,start) (update_gtf.c:41)");
    return LRB_OK;
}

__global__ void invert_mask_kernel(const uint8_t *a, uint8_t *o, int64_t n)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) o[i] = !a[i];
}

// empty-chain / sortedness check for unique (classification does it for update)
__global__ void rows_check_kernel(DRows rows, uint32_t *err, int check_sorted)
{
    int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= rows.n) return;
    if (rows.ex_n[r] == 0) atomicOr(err, 2u);
    if (check_sorted && r > 0) {                     // bit 2: not sorted by (tid,start) -- allowed for unique (a concatenation of samples)
        const int pt = rows.tid[r - 1], ps = rows.start[r - 1], ct = rows.tid[r], cs = rows.start[r];
        if (pt > ct || (pt == ct && ps > cs)) atomicOr(err, 4u);
    }
}

int lrb_update_run(lrb_ctx *c, const lrb_update_params *up)
{
    if (!c || !up) return LRB_E_ARG;
    if (!c->have_exons) return fail(c, LRB_E_ARG, "lrb_update_run: exon chains missing (run lrb_exon_run / lrb_pipeline_run / lrb_chains_upload)");
    CK(cudaSetDevice(c->device));
    int64_t l0 = total_launches();
    DRows &rows = *c->cur; const int64_t n = rows.n; const size_t nn = (size_t)std::max<int64_t>(n, 1);
    int rc;
    memset(c->summary, 0, sizeof c->summary); c->n_bed = 0; c->last_up = *up; c->n_known = c->n_unrecog = 0;
    NEED(c->u_cls, nn * 4); NEED(c->u_ref, nn * 4); NEED(c->u_nnovel, nn * 4);
    NEED(c->u_mk, nn); NEED(c->u_known, nn * 4); NEED(c->u_unrecog, nn * 4); NEED(c->u_ck, nn);
    NEED(c->y_counts, 64); NEED(c->y_nelem, 64);       // y_nelem: element count, then (at +16) the six tickets of sum_scans_kernel
    if ((rc = ensure_tiles(c, std::max<int64_t>(n, c->ex.n)))) return rc;
    CK(cudaMemsetAsync(d_err(c), 0, 4, c->st));
    tick(c, 0);
    // ---- classification (+ SJ support)
    ClassArgs ca{};
    ca.rows = rows; ca.ex = c->ex; ca.anno = c->anno; ca.sj = c->sj; ca.up = *up;
    ca.row_nonmono = c->have_batch ? nullptr : c->r_nonmono.as<uint8_t>();
    ca.cls = c->u_cls.as<uint32_t>(); ca.ref = c->u_ref.as<int32_t>(); ca.n_novel = c->u_nnovel.as<uint32_t>(); ca.err_flags = d_err(c);
    launch_classify(ca, c->u_mk.as<uint8_t>(), c->st);     // u_mk: scratch for the slow-row mask
    CK(cudaGetLastError());
    tick(c, 1);

    // ---- lists, the updated_T fold, the class folds and the element count of the summary sets: one stream of launches,
    // every count stays on the device; ONE copy brings them all back.  novel_T is sized optimistically (pieces are rare);
    // if it turns out too small the pass is repeated once with the exact size.
    uint64_t *T = d_totals(c);
    int64_t cap = up->split_trans ? n + n / 8 + 1024 : n;
    cap = std::max<int64_t>(cap, std::min<int64_t>(c->novel_cap_hint, n + c->ex.n / 2 + 1));
    if (c->test_small_novel_cap) cap = std::max<int64_t>(1, n / 4);   // test hook: forces the undersized-list retry
    // A split piece scans the WHOLE of updated_T in the reference (its tid/start/end are 0: update_gtf.c:148 never stops it), so it
    // can be absorbed by a chain on another chromosome.  The locus-parallel fold cannot see that; the set kernels probe for the
    // necessary condition (a junction of a piece that also exists on another chromosome, CNT_XLOCUS) and the fold is then replayed
    // once more as ONE locus (merge_fold_kernel: the exact back-scan, entry by entry).  With -d > 0 junctions match approximately and
    // the exact-key probe proves nothing: any surviving piece sends the fold to the replay.
    // With the default exact matching (-d 0, no -D) that is settled in rounds on the device (xl_* kernels in lrb_update.cu): exact and
    // cheap.  The one-locus replay remains for -d > 0 / -D, where check_iden depends on the entries' moving ends.
    const bool xl_exact = up->split_trans != 0 && up->ss_dis == 0 && up->end_dis == 0x7fffffff && !c->force_single_fold;
    const bool detect = up->split_trans != 0 && !xl_exact, run_sets = up->want_summary || detect;
    bool single = c->force_single_fold;
    c->n_xlocus_pieces = 0;
    for (int fold_pass = 0; fold_pass < 2; ++fold_pass) {
    SummaryArgs sa{};
    int64_t n_novel = 0, nu = 0; uint64_t n_elem = 0, n_exon_elem = 0;
    uint32_t cnt16[16];
    for (int attempt = 0; n > 0; ++attempt) {
        if ((rc = setup_list(c, c->novel, c->n_row, c->n_lo, c->n_cnt, c->n_piece, cap))) return rc;
        c->novel.cap = cap;
        CK(cudaMemsetAsync(c->y_counts.p, 0, 64, c->st)); CK(cudaMemsetAsync(c->y_nelem.p, 0, 16, c->st));
        ListArgs la{};
        la.rows = rows; la.ex = c->ex; la.up = *up; la.cls = ca.cls; la.n_novel = ca.n_novel; la.novel = c->novel;
        la.known = c->u_known.as<uint32_t>(); la.unrecog = c->u_unrecog.as<uint32_t>();
        la.kls = up->want_summary ? c->u_ck.as<uint8_t>() : nullptr; la.class_n = c->y_counts.as<uint32_t>() + 12;
        la.tile_state = c->tile_state.as<uint64_t>(); la.ticket = d_ticket(c); la.totals = T + T_NOVEL;
        launch_build_lists(la, c->st);
        CK(cudaGetLastError());
        if (attempt == 0 && fold_pass == 0) tick(c, 2);
        if (up->want_summary) {
            // class counts + uniq_* folds over bam_T (update_gtf.c:501-528): the four classes partition the rows; they are
            // folded in ONE pass over all rows, each candidate seeing only the entries of its own class.  Independent of the
            // updated_T fold and of the sets: forked onto the side stream, joined in front of the round trip below.
            const bool side = c->side_stream;
            cudaStream_t s2 = side ? c->st2 : c->st;
            if ((rc = setup_list(c, c->tmp_list, c->t_row, c->t_lo, c->t_cnt, c->t_piece, n))) return rc;
            if (side) { CK(cudaEventRecord(c->ev_fork, c->st)); CK(cudaStreamWaitEvent(c->st2, c->ev_fork, 0)); }
            launch_rows_as_list(rows, nullptr, n, c->tmp_list, s2);
            if ((rc = run_merge_async(c, c->mg2, c->tmp_list, n, nullptr, *up, T + T_LOCI2, c->u_ck.as<uint8_t>(), c->y_counts.as<uint32_t>() + 8, false, side))) return rc;
            if (side) CK(cudaEventRecord(c->ev_join, c->st2));
        }
        // updated_T = merge fold over novel_T (update_gtf.c:949,956)
        if (xl_exact) { NEED(c->xl_forced, (size_t)std::max<int64_t>(cap, 1)); CK(cudaMemsetAsync(c->xl_forced.p, 0, (size_t)std::max<int64_t>(cap, 1), c->st)); }
        if ((rc = run_merge_async(c, c->mg, c->novel, cap, T + T_NOVEL, *up, T + T_LOCI, nullptr, nullptr, true, false, single, xl_exact))) return rc;
        if (attempt == 0 && fold_pass == 0) tick(c, 3);
        if (run_sets) {
            // sets over updated_T: element count (sizes the hash table)
            const size_t capn = (size_t)std::max<int64_t>(cap, 1);
            NEED(c->y_barcnt, capn * 16); NEED(c->y_barseg, capn * 16); NEED(c->y_genebar, capn * 8); NEED(c->y_bedcnt, capn * 4); NEED(c->y_bedoff, capn * 4);
            sa = SummaryArgs{}; sa.sets = up->want_summary ? SUM_ALL : 0; sa.probe = detect ? 1 : 0;
            sa.rows = rows; sa.ex = c->ex; sa.list = c->novel; sa.n_upd = cap; sa.n_upd_dev = T + T_UPD;
            sa.upd = merged_view(c->mg.o_cand, c->mg.o_cov, c->mg.o_tid, c->mg.o_start, c->mg.o_end, c->mg.o_fs, c->mg.o_le, cap);
            sa.ref = c->u_ref.as<int32_t>(); sa.anno_gene = c->anno.gene;
            sa.bar_cnt = c->y_barcnt.as<uint32_t>(); sa.bar_seg = c->y_barseg.as<uint32_t>(); sa.gene_bar = c->y_genebar.as<uint64_t>();
            sa.bed_cnt = c->y_bedcnt.as<uint32_t>(); sa.bed_off = c->y_bedoff.as<uint32_t>(); sa.counts = c->y_counts.as<uint32_t>();
            launch_summary_count(sa, c->y_nelem.as<unsigned long long>(), c->st);
            CK(cudaGetLastError());
        }
        // ---- the one round trip: totals, error flags, class sizes, element count.  The class folds keep running on the
        // side stream underneath the summary sets; they are joined in front of the last copy of this stage.
        uint8_t *hp = (uint8_t *)c->h_scalars.p;
        CK(cudaMemcpyAsync(hp, c->scalars.p, T_SLOTS * 8 + 8, cudaMemcpyDeviceToHost, c->st));
        CK(cudaMemcpyAsync(hp + 320, c->y_counts.p, 64, cudaMemcpyDeviceToHost, c->st));
        CK(cudaMemcpyAsync(hp + 384, c->y_nelem.p, 16, cudaMemcpyDeviceToHost, c->st));
        if (xl_exact) CK(cudaMemcpyAsync(hp + 400, c->xl_cnt.p, XL_NCNT * 4, cudaMemcpyDeviceToHost, c->st));
        CK(cudaStreamSynchronize(c->st));
        uint64_t t[T_SLOTS]; uint32_t e[2], xc[XL_NCNT] = {0};
        memcpy(t, hp, sizeof t); memcpy(e, hp + T_SLOTS * 8, 8); memcpy(cnt16, hp + 320, 64); memcpy(&n_elem, hp + 384, 8); memcpy(&n_exon_elem, hp + 392, 8);
        if (xl_exact) memcpy(xc, hp + 400, sizeof xc);
        if (e[1] & 3u) {
            // the class folds may still be running on the side stream: they must be over before the caller can start another
            // stage on this context (they share its buffers and counters)
            if (up->want_summary && c->side_stream) cudaStreamSynchronize(c->st2);
            if (e[1] & 2u) return fail(c, LRB_E_UNMAPPED, "unmapped record / empty exon chain in update/unique input (the reference aborts here, bam2gtf.c:95-100)");
            return fail(c, LRB_E_UNSORTED, "reads are not sorted by (tid,start) (update_gtf.c:41)");
        }
        n_novel = (int64_t)t[T_NOVEL]; c->n_known = (int64_t)t[T_KNOWN]; c->n_unrecog = (int64_t)t[T_UNREC];
        if (n_novel > cap) {                          // novel_T did not fit: once more with the exact size
            if (attempt >= 1) return fail(c, LRB_E_CUDA, "novel_T size changed between passes");
            cap = n_novel; c->novel_cap_hint = n_novel + n_novel / 16;
            if (up->want_summary && c->side_stream) CK(cudaStreamWaitEvent(c->st, c->ev_join, 0));   // the counters are reset below
            continue;
        }
        c->mg.n_loci = (int64_t)t[T_LOCI]; c->mg.n_out = nu = (int64_t)t[T_UPD];
        if (xl_exact && (xc[XL_OVERFLOW] || xc[XL_CHANGED])) {
            // some piece is absorbed by an entry on another chromosome: fold again with the marks until they are stable (round k fixes
            // the k-th affected piece at the latest), then let the absorbing entries take cov / ends, compact and count again
            int rounds = 0;
            while (!xc[XL_OVERFLOW] && xc[XL_CHANGED] && rounds < 32) {
                ++rounds;
                launch_xlocus_reset(c->mg.args, c->xl, c->st);
                launch_merge_fold(c->mg.args, c->st);
                launch_xlocus_detect(c->mg.args, c->xl, c->st);
                CK(cudaMemcpyAsync(hp + 400, c->xl_cnt.p, XL_NCNT * 4, cudaMemcpyDeviceToHost, c->st));
                CK(cudaStreamSynchronize(c->st));
                memcpy(xc, hp + 400, sizeof xc);
            }
            if (xc[XL_OVERFLOW] || xc[XL_CHANGED]) {                 // more hits than the lists hold, or no fixed point in 32 rounds: the exact replay
                if (up->want_summary && c->side_stream) CK(cudaStreamWaitEvent(c->st, c->ev_join, 0));
                single = true; c->n_xlocus_replays++; n_novel = -1; break;
            }
            c->n_xlocus_pieces = xc[XL_NFORCED];
            launch_xlocus_apply(c->mg.args, c->xl, c->st);
            CK(cudaMemsetAsync(T + T_UPD, 0, 8, c->st));
            launch_merge_finish(c->mg.args, c->st);
            if (run_sets) { CK(cudaMemsetAsync(c->y_nelem.p, 0, 16, c->st)); launch_summary_count(sa, c->y_nelem.as<unsigned long long>(), c->st); }
            CK(cudaGetLastError());
            CK(cudaMemcpyAsync(hp, c->scalars.p, T_SLOTS * 8 + 8, cudaMemcpyDeviceToHost, c->st));
            CK(cudaMemcpyAsync(hp + 384, c->y_nelem.p, 16, cudaMemcpyDeviceToHost, c->st));
            CK(cudaStreamSynchronize(c->st));
            memcpy(t, hp, sizeof t); memcpy(&n_elem, hp + 384, 8); memcpy(&n_exon_elem, hp + 392, 8);
            c->mg.n_out = nu = (int64_t)t[T_UPD];
        }
        break;
    }
    if (n_novel < 0) continue;                                        // the marks did not settle: once more as ONE locus
    if (n == 0) { if ((rc = setup_list(c, c->novel, c->n_row, c->n_lo, c->n_cnt, c->n_piece, 0))) return rc; c->mg.n_out = c->mg.n_loci = 0; tick(c, 2); tick(c, 3); }
    c->novel.n = n_novel; c->novel.cap = std::max<int64_t>(cap, 0);
    if (c->timing && n && fold_pass == 0) { CK(cudaStreamSynchronize(c->st)); cudaEventElapsedTime(&c->ms[LRB_T_K_FOLD], c->ev[10], c->ev[11]); }

    // ---- summary sets (print_trans_summary, update_gtf.c:421-587) and / or the probe for pieces that meet another chromosome
    if (!(run_sets && n > 0)) break;
    int32_t *s = c->summary;
    const int cnt_idx[4] = {LRB_S_KNOWN_TRANS, LRB_S_NOVEL_RELIABLE, LRB_S_NOVEL_UNRELIABLE, LRB_S_UNRECOG};
    const int uniq_idx[4] = {LRB_S_UNIQ_KNOWN, LRB_S_UNIQ_RELIABLE, LRB_S_UNIQ_UNRELIABLE, LRB_S_UNIQ_UNRECOG};
    const int64_t n_known_reads = up->want_summary ? (int64_t)cnt16[12] : 0;
    sa.n_upd = nu; sa.n_upd_dev = nullptr; sa.upd.n = nu;
    // distinct table keys: one per exon element, up to two per gene / site / junction element (the tid-0 phase and the
    // segment phase key them differently), one per known read; the table stays below that bound's next power of two
    // (load <= 0.8 in the worst case, ~0.4 on the bench shape; the slot index is a multiply-high, so no power of two is needed)
    const uint64_t worst = 2 * n_elem - std::min(n_exon_elem, n_elem) + (uint64_t)n_known_reads;
    const uint64_t capn = worst + worst / 4 + 1024;
    NEED(c->h_khi, capn * sizeof(HashSlot));
    CK(cudaMemsetAsync(c->h_khi.p, 0xFF, capn * sizeof(HashSlot), c->st));
    sa.tab.cap = capn; sa.tab.slots = c->h_khi.as<HashSlot>();
    // BED rows are the first occurrences of the exon set: at most one per counted element
    const size_t nb = (size_t)std::max<uint64_t>(n_elem, 1);
    NEED(c->bd_tid, nb * 4); NEED(c->bd_s, nb * 4); NEED(c->bd_e, nb * 4); NEED(c->bd_sc, nb * 4); NEED(c->bd_ty, nb); NEED(c->bd_rv, nb);
    sa.bed_tid = c->bd_tid.as<int32_t>(); sa.bed_start = c->bd_s.as<int32_t>(); sa.bed_end = c->bd_e.as<int32_t>(); sa.bed_score = c->bd_sc.as<int32_t>();
    sa.bed_type = c->bd_ty.as<uint8_t>(); sa.bed_rev = c->bd_rv.as<uint8_t>();
    if (c->want_kg_pairs && up->want_summary) { NEED(c->kg_pairs, (size_t)std::max<int64_t>(n_known_reads, 1) * 8); sa.kg_pairs = c->kg_pairs.as<int2>(); }
    CK(cudaMemsetAsync(c->y_counts.p, 0, 32, c->st));
    launch_summary_sets(sa, ca.cls, n, c->tile_state.as<uint64_t>(), (uint32_t *)(c->y_nelem.as<uint8_t>() + 16), T + T_BED, c->st,
                        c->sum_split ? c->st3 : c->st, c->ev_fork3, c->ev_join3);
    CK(cudaGetLastError());
    if (up->want_summary && c->side_stream) CK(cudaStreamWaitEvent(c->st, c->ev_join, 0));
    uint8_t *hp = (uint8_t *)c->h_scalars.p;
    CK(cudaMemcpyAsync(hp, T + T_BED, 8, cudaMemcpyDeviceToHost, c->st));
    CK(cudaMemcpyAsync(hp + 64, c->y_counts.p, 48, cudaMemcpyDeviceToHost, c->st));
    CK(cudaStreamSynchronize(c->st));
    uint64_t nbed; uint32_t cnt[12];
    memcpy(&nbed, hp, 8); memcpy(cnt, hp + 64, 48);
    c->xlocus_seen = detect && (cnt[CNT_XLOCUS] != 0 || (up->ss_dis > 0 && cnt[CNT_PARTIAL] != 0));
    if (c->xlocus_seen && !single) { single = true; c->n_xlocus_replays++; continue; }       // replay the fold as one locus
    if (up->want_summary) {
        for (int k = 0; k < 4; ++k) s[cnt_idx[k]] = (int32_t)cnt16[12 + k];
        s[LRB_S_NOVEL_BAM] = s[LRB_S_NOVEL_RELIABLE] + s[LRB_S_NOVEL_UNRELIABLE];
        for (int k = 0; k < 4; ++k) s[uniq_idx[k]] = (int32_t)cnt[8 + k];
        c->n_bed = nu ? (int64_t)nbed : 0;
        s[LRB_S_UPD_GENES] = (int32_t)cnt[CNT_G]; s[LRB_S_NOVEL_TRANS] = (int32_t)nu; s[LRB_S_NOVEL_PARTIAL] = (int32_t)cnt[CNT_PARTIAL];
        s[LRB_S_NOVEL_FULL] = (int32_t)nu - (int32_t)cnt[CNT_PARTIAL]; s[LRB_S_NOVEL_EXONS] = (int32_t)cnt[CNT_E]; s[LRB_S_NOVEL_SITES] = (int32_t)(cnt[CNT_D] + cnt[CNT_A]);
        s[LRB_S_NOVEL_JUNC] = (int32_t)cnt[CNT_J]; s[LRB_S_KNOWN_GENES] = (int32_t)cnt[CNT_KG];
        c->n_kg_pairs = (int64_t)cnt[CNT_KG];
    }
    break;
    }   // fold_pass
    tick(c, 4);
    c->have_update = true; c->launches_last = total_launches() - l0;
    if (c->timing) {
        CK(cudaStreamSynchronize(c->st));
        cudaEventElapsedTime(&c->ms[LRB_T_CLASSIFY], c->ev[0], c->ev[1]);
        float lists = 0; cudaEventElapsedTime(&lists, c->ev[1], c->ev[2]);
        cudaEventElapsedTime(&c->ms[LRB_T_MERGE], c->ev[2], c->ev[3]); c->ms[LRB_T_MERGE] += lists;
        cudaEventElapsedTime(&c->ms[LRB_T_SUMMARY], c->ev[3], c->ev[4]);
    }
    return LRB_OK;
}

int lrb_unique_run(lrb_ctx *c, const lrb_update_params *up)
{
    if (!c || !up) return LRB_E_ARG;
    if (!c->have_exons) return fail(c, LRB_E_ARG, "lrb_unique_run: exon chains missing");
    CK(cudaSetDevice(c->device));
    int64_t l0 = total_launches();
    DRows &rows = *c->cur; const int64_t n = rows.n;
    int rc;
    CK(cudaMemsetAsync(d_err(c), 0, 4, c->st));
    tick(c, 0);
    if (n) { rows_check_kernel<<<(unsigned)((n + 255) / 256), 256, 0, c->st>>>(rows, d_err(c), 1); CK(cudaGetLastError()); }
    uint32_t flags = 0;
    if ((rc = check_err_flags(c, &flags))) return rc;
    if ((rc = setup_list(c, c->tmp_list, c->t_row, c->t_lo, c->t_cnt, c->t_piece, n))) return rc;
    launch_rows_as_list(rows, nullptr, n, c->tmp_list, c->st);
    // unique-gtf's input may be a concatenation of sorted samples (Snakefile:189-192).  The locus cuts (App. B.3) are only
    // proven for (tid,start)-sorted streams, so an unsorted list is folded as ONE locus: the back-scan of merge_trans is
    // replayed entry by entry (merge_fold_kernel), early-termination quirks included.
    if ((rc = run_merge(c, c->mg, c->tmp_list, n, *up, (flags & 4u) != 0))) return rc;
    // shared_T = rows the fold absorbed: complement of the alive mask (kept in mg.dropped); mg.head is free again
    NEED(c->q_shared, (size_t)std::max<int64_t>(n, 1) * 4);
    c->n_shared = 0;
    if (n) {
        invert_mask_kernel<<<(unsigned)((n + 255) / 256), 256, 0, c->st>>>(c->mg.dropped.as<uint8_t>(), c->mg.head.as<uint8_t>(), n);
        launch_compact_mask(c->mg.head.as<uint8_t>(), n, nullptr, c->q_shared.as<uint32_t>(), nullptr, c->tile_state.as<uint64_t>(), d_ticket(c), d_totals(c), c->st);
        CK(cudaGetLastError());
        uint64_t t; if ((rc = read_totals(c, &t, 1))) return rc;
        c->n_shared = (int64_t)t;
    }
    tick(c, 1);
    c->have_unique = true; c->launches_last = total_launches() - l0;
    if (c->timing) { CK(cudaStreamSynchronize(c->st)); cudaEventElapsedTime(&c->ms[LRB_T_MERGE], c->ev[0], c->ev[1]); }
    return LRB_OK;
}

// Coordinate sort of the current rows on the device: what `samtools sort` does between `lr2rmats filter` and
// `lr2rmats update-gtf` (Snakefile:90), without leaving HBM.  Stable LSD radix sort on tid << 32 | (pos + 1) << 1 | FLAG 0x10.
int lrb_rows_sort(lrb_ctx *c)
{
    if (!c) return LRB_E_ARG;
    if (!c->have_exons) return fail(c, LRB_E_ARG, "lrb_rows_sort: exon chains missing (run lrb_exon_run / lrb_pipeline_run / lrb_chains_upload)");
    CK(cudaSetDevice(c->device));
    int64_t l0 = total_launches();
    if (c->cur == &c->rows3) return LRB_OK;                           // sorted already (every upload / stage run resets cur)
    DRows &rows = *c->cur; const int64_t n = rows.n; int rc;
    if (n == 0) return LRB_OK;
    if (n >= (int64_t)1 << 32) return fail(c, LRB_E_ARG, "lrb_rows_sort: more than 2^32 rows");
    const size_t nn = (size_t)n;
    NEED(c->s_key0, nn * 8); NEED(c->s_key1, nn * 8); NEED(c->s_idx0, nn * 4); NEED(c->s_idx1, nn * 4);
    NEED(c->s_hist, (size_t)256 * (size_t)lrbk::sort_tiles(n) * 4);
    if ((rc = setup_rows(c, c->rows3, c->s_read, c->s_rtid, c->s_rs, c->s_re, c->s_rev, c->s_beg, c->s_n, n))) return rc;
    unsigned long long *d_max = (unsigned long long *)(d_totals(c) + 3);
    CK(cudaMemsetAsync(d_max, 0, 8, c->st));
    launch_sort_keys(rows, c->have_batch ? c->b.flag : nullptr, c->s_key0.as<uint64_t>(), c->s_idx0.as<uint32_t>(), d_max, c->st);
    CK(cudaGetLastError());
    uint64_t t[4];
    if ((rc = read_totals(c, t, 4))) return rc;
    int bits = 0; for (uint64_t m = t[3]; m; m >>= 1) ++bits;
    uint64_t *k[2] = {c->s_key0.as<uint64_t>(), c->s_key1.as<uint64_t>()}; uint32_t *v[2] = {c->s_idx0.as<uint32_t>(), c->s_idx1.as<uint32_t>()};
    int cur = 0;
    for (int shift = 0; shift < bits; shift += 8) {                  // only the digits the largest key has
        launch_sort_pass(k[cur], v[cur], k[cur ^ 1], v[cur ^ 1], n, shift, c->s_hist.as<uint32_t>(), c->st);
        cur ^= 1;
    }
    c->rows3.n = n;
    launch_rows_permute(rows, c->rows3, v[cur], c->st);
    CK(cudaGetLastError());
    c->cur = &c->rows3; c->rows_compact = false; c->have_update = c->have_unique = false;
    c->launches_last = total_launches() - l0;
    return LRB_OK;
}

int lrb_sync(lrb_ctx *c)
{
    if (!c) return LRB_E_ARG;
    CK(cudaSetDevice(c->device));
    CK(cudaStreamSynchronize(c->st));
    return LRB_OK;
}

// ------------------------------------------------------------------------------------------------- fetches
int lrb_filter_fetch(lrb_ctx *c, lrb_filter_result *out)
{
    if (!c || !out) return LRB_E_ARG;
    if (!c->have_filter) return fail(c, LRB_E_ARG, "lrb_filter_fetch: filter stage has not run");
    CK(cudaSetDevice(c->device));
    const size_t n = (size_t)c->b.n; int rc;
    if ((rc = d2h(c, c->p[0], c->f_pass.as<uint8_t>(), n))) return rc;
    if ((rc = d2h(c, c->p[1], c->f_score.as<int32_t>(), n))) return rc;
    if ((rc = d2h(c, c->p[2], c->f_intron.as<int32_t>(), n))) return rc;
    if ((rc = d2h(c, c->p[3], c->f_keep_idx.as<uint32_t>(), (size_t)c->n_keep))) return rc;
    CK(cudaStreamSynchronize(c->st));
    out->n = c->b.n; out->pass = c->p[0].as<uint8_t>(); out->score = c->p[1].as<int32_t>(); out->intron_n = c->p[2].as<int32_t>();
    out->n_keep = c->n_keep; out->keep_idx = c->p[3].as<uint32_t>();
    return LRB_OK;
}

int lrb_filter_fetch_keep(lrb_ctx *c, int64_t *n_keep, const uint32_t **keep_idx)
{
    if (!c || !n_keep || !keep_idx) return LRB_E_ARG;
    if (!c->have_filter) return fail(c, LRB_E_ARG, "lrb_filter_fetch_keep: filter stage has not run");
    CK(cudaSetDevice(c->device));
    int rc;
    if ((rc = d2h(c, c->p[3], c->f_keep_idx.as<uint32_t>(), (size_t)c->n_keep))) return rc;
    CK(cudaStreamSynchronize(c->st));
    *n_keep = c->n_keep; *keep_idx = c->p[3].as<uint32_t>();
    return LRB_OK;
}

// compact copy of the exon data of the current rows (rows may reference a sparse subset of the pool after the fused pass)
__global__ void exon_gather_kernel(DRows rows, DExons ex, const uint32_t *__restrict__ off, int32_t *es, int32_t *ee, uint8_t *fl, int with_flags)
{
    int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= rows.n) return;
    uint32_t b = rows.ex_beg[r], n = rows.ex_n[r], o = off[r];
    for (uint32_t j = 0; j < n; ++j) { es[o + j] = ex.es[b + j]; ee[o + j] = ex.ee[b + j]; if (with_flags) fl[o + j] = ex.flag[b + j]; }
}

static int fetch_exons(lrb_ctx *c, lrb_exon_result *out, bool with_flags, const uint8_t **flags_out, int pbase)
{
    DRows &rows = *c->cur; const int64_t n = rows.n; int rc;
    if ((rc = d2h(c, c->p[pbase + 0], rows.read_idx, (size_t)n))) return rc;
    if ((rc = d2h(c, c->p[pbase + 1], rows.tid, (size_t)n))) return rc;
    if ((rc = d2h(c, c->p[pbase + 2], rows.is_rev, (size_t)n))) return rc;
    int64_t ne = 0;
    NEEDP(c->p[pbase + 3], ((size_t)n + 1) * 4);
    if (c->rows_compact) {
        ne = c->ex.n;
        if (n) CK(cudaMemcpyAsync(c->p[pbase + 3].p, rows.ex_beg, (size_t)n * 4, cudaMemcpyDeviceToHost, c->st));
        if ((rc = d2h(c, c->p[pbase + 4], c->ex.es, (size_t)ne))) return rc;
        if ((rc = d2h(c, c->p[pbase + 5], c->ex.ee, (size_t)ne))) return rc;
        if (with_flags && (rc = d2h(c, c->p[pbase + 6], c->ex.flag, (size_t)ne))) return rc;
        CK(cudaStreamSynchronize(c->st));
        c->p[pbase + 3].as<uint32_t>()[n] = (uint32_t)ne;
    } else {
        // exclusive scan of ex_n -> offsets, then gather into scratch (the summary hash buffers are free at fetch time)
        if ((rc = ensure_tiles(c, n))) return rc;
        NEED(c->u_sub, ((size_t)n + 1) * 4);
        launch_scan_sum_u32(rows.ex_n, c->u_sub.as<uint32_t>(), n, c->tile_state.as<uint64_t>(), d_ticket(c), d_totals(c), c->st);
        uint64_t t; if ((rc = read_totals(c, &t, 1))) return rc;
        ne = n ? (int64_t)t : 0;
        const size_t k = (size_t)std::max<int64_t>(ne, 1);
        NEED(c->y_barcnt, k * 4); NEED(c->y_barseg, k * 4); NEED(c->y_bedcnt, k);
        if (n) {
            exon_gather_kernel<<<(unsigned)((n + 255) / 256), 256, 0, c->st>>>(rows, c->ex, c->u_sub.as<uint32_t>(), c->y_barcnt.as<int32_t>(),
                                                                                 c->y_barseg.as<int32_t>(), c->y_bedcnt.as<uint8_t>(), with_flags ? 1 : 0);
            CK(cudaGetLastError());
            CK(cudaMemcpyAsync(c->p[pbase + 3].p, c->u_sub.p, (size_t)n * 4, cudaMemcpyDeviceToHost, c->st));
        }
        if ((rc = d2h(c, c->p[pbase + 4], c->y_barcnt.as<int32_t>(), (size_t)ne))) return rc;
        if ((rc = d2h(c, c->p[pbase + 5], c->y_barseg.as<int32_t>(), (size_t)ne))) return rc;
        if (with_flags && (rc = d2h(c, c->p[pbase + 6], c->y_bedcnt.as<uint8_t>(), (size_t)ne))) return rc;
        CK(cudaStreamSynchronize(c->st));
        c->p[pbase + 3].as<uint32_t>()[n] = (uint32_t)ne;
    }
    out->n_reads = n; out->read_idx = c->have_batch ? c->p[pbase + 0].as<uint32_t>() : nullptr; out->tid = c->p[pbase + 1].as<int32_t>();
    out->is_rev = c->p[pbase + 2].as<uint8_t>(); out->exon_off = c->p[pbase + 3].as<uint32_t>();
    out->exon_start = c->p[pbase + 4].as<int32_t>(); out->exon_end = c->p[pbase + 5].as<int32_t>();
    if (flags_out) *flags_out = with_flags ? c->p[pbase + 6].as<uint8_t>() : nullptr;
    return LRB_OK;
}

int lrb_exon_fetch(lrb_ctx *c, lrb_exon_result *out)
{
    if (!c || !out) return LRB_E_ARG;
    if (!c->have_exons) return fail(c, LRB_E_ARG, "lrb_exon_fetch: exon stage has not run");
    CK(cudaSetDevice(c->device));
    return fetch_exons(c, out, false, nullptr, 4);
}

static int fetch_merged(lrb_ctx *c, MergeBufs &m, lrb_merged_list *o, int pbase)
{
    const size_t n = (size_t)m.n_out; int rc;
    Buf *src[] = {&m.o_cand, &m.o_cov, &m.o_tid, &m.o_start, &m.o_end, &m.o_fs, &m.o_le};
    for (int i = 0; i < 7; ++i) if ((rc = d2h(c, c->p[pbase + i], src[i]->as<uint32_t>(), n))) return rc;
    o->n = m.n_out; o->cand = c->p[pbase].as<uint32_t>(); o->cov = c->p[pbase + 1].as<int32_t>(); o->t_tid = c->p[pbase + 2].as<int32_t>();
    o->t_start = c->p[pbase + 3].as<int32_t>(); o->t_end = c->p[pbase + 4].as<int32_t>(); o->first_start = c->p[pbase + 5].as<int32_t>();
    o->last_end = c->p[pbase + 6].as<int32_t>();
    return LRB_OK;
}

int lrb_update_fetch(lrb_ctx *c, lrb_update_result *out)
{
    if (!c || !out) return LRB_E_ARG;
    if (!c->have_update) return fail(c, LRB_E_ARG, "lrb_update_fetch: update stage has not run");
    CK(cudaSetDevice(c->device));
    memset(out, 0, sizeof *out);
    int rc; const int64_t n = c->cur->n;
    if ((rc = fetch_exons(c, &out->ex, true, &out->exon_flag, 4))) return rc;
    if ((rc = d2h(c, c->p[11], c->u_cls.as<uint32_t>(), (size_t)n))) return rc;
    if ((rc = d2h(c, c->p[12], c->u_ref.as<int32_t>(), (size_t)n))) return rc;
    if ((rc = d2h(c, c->p[13], c->u_known.as<uint32_t>(), (size_t)c->n_known))) return rc;
    if ((rc = d2h(c, c->p[14], c->u_unrecog.as<uint32_t>(), (size_t)c->n_unrecog))) return rc;
    if ((rc = d2h(c, c->p[15], c->novel.row, (size_t)c->novel.n))) return rc;
    if ((rc = d2h(c, c->p[16], c->novel.lo, (size_t)c->novel.n))) return rc;
    if ((rc = d2h(c, c->p[17], c->novel.cnt, (size_t)c->novel.n))) return rc;
    if ((rc = d2h(c, c->p[18], c->novel.piece, (size_t)c->novel.n))) return rc;
    if ((rc = fetch_merged(c, c->mg, &out->updated, 19))) return rc;
    if ((rc = d2h(c, c->p[26], c->bd_tid.as<int32_t>(), (size_t)c->n_bed))) return rc;
    if ((rc = d2h(c, c->p[27], c->bd_s.as<int32_t>(), (size_t)c->n_bed))) return rc;
    if ((rc = d2h(c, c->p[28], c->bd_e.as<int32_t>(), (size_t)c->n_bed))) return rc;
    if ((rc = d2h(c, c->p[29], c->bd_sc.as<int32_t>(), (size_t)c->n_bed))) return rc;
    if ((rc = d2h(c, c->p[30], c->bd_ty.as<uint8_t>(), (size_t)c->n_bed))) return rc;
    if ((rc = d2h(c, c->p[31], c->bd_rv.as<uint8_t>(), (size_t)c->n_bed))) return rc;
    CK(cudaStreamSynchronize(c->st));
    out->cls = c->p[11].as<uint32_t>(); out->ref_anno = c->p[12].as<int32_t>();
    out->n_known = c->n_known; out->known_idx = c->p[13].as<uint32_t>(); out->n_unrecog = c->n_unrecog; out->unrecog_idx = c->p[14].as<uint32_t>();
    out->novel.n = c->novel.n; out->novel.read = c->p[15].as<uint32_t>(); out->novel.exon_lo = c->p[16].as<uint32_t>();
    out->novel.exon_n = c->p[17].as<uint32_t>(); out->novel.piece = c->p[18].as<int32_t>();
    memcpy(out->summary, c->summary, sizeof c->summary);
    out->bed.n = c->n_bed; out->bed.tid = c->p[26].as<int32_t>(); out->bed.start = c->p[27].as<int32_t>(); out->bed.end = c->p[28].as<int32_t>();
    out->bed.score = c->p[29].as<int32_t>(); out->bed.type = c->p[30].as<uint8_t>(); out->bed.is_rev = c->p[31].as<uint8_t>();
    return LRB_OK;
}

// ---- updated_T as a self-contained table (what print_read_trans needs), gathered on the device
struct TabArgs {
    DRows rows; DExons ex; DTransList list; DMerged upd; const int32_t *ref; int rows_have_read_idx;
    uint32_t *name_idx; int32_t *piece, *t_tid, *t_start, *t_end, *e_tid, *cov, *ref_out; uint8_t *t_rev, *e_rev; uint32_t *cnt;
    const uint32_t *off; int32_t *es, *ee; uint8_t *fl;
};
__global__ void tab_rows_kernel(TabArgs a)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= a.upd.n) return;
    const uint32_t cd = a.upd.cand[i], row = a.list.row[cd];
    const int piece = a.list.piece[cd];
    const uint8_t rev = a.rows.is_rev[row];
    a.name_idx[i] = a.rows_have_read_idx ? a.rows.read_idx[row] : row;
    a.piece[i] = piece; a.t_tid[i] = a.upd.tid[i]; a.t_start[i] = a.upd.start[i]; a.t_end[i] = a.upd.end[i];
    a.t_rev[i] = piece >= 0 ? 0 : rev; a.e_tid[i] = a.rows.tid[row]; a.e_rev[i] = rev; a.cov[i] = a.upd.cov[i]; a.ref_out[i] = a.ref[row];
    a.cnt[i] = a.list.cnt[cd];
}
__global__ void tab_exons_kernel(TabArgs a)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= a.upd.n) return;
    const uint32_t cd = a.upd.cand[i], row = a.list.row[cd];
    const uint32_t gb = a.rows.ex_beg[row] + a.list.lo[cd], o = a.off[i];
    const int n = (int)a.list.cnt[cd];
    for (int j = 0; j < n; ++j) {
        a.es[o + j] = j == 0 ? a.upd.fs[i] : a.ex.es[gb + j];
        a.ee[o + j] = j == n - 1 ? a.upd.le[i] : a.ex.ee[gb + j];
        if (a.fl) a.fl[o + j] = a.ex.flag[gb + j];
    }
}

}  // extern "C" (reopened below)

// updated_T as a self-contained table in device memory (c->tb_*): rows + exon offsets + exon pools (+ the exon flag bytes for the
// gather root of a multi-GPU run).  One host round trip (the exon total).
int lrbk::build_update_table(lrb_ctx *c, int64_t *n_exon_out, bool with_flags)
{
    int rc;
    const int64_t nu = c->mg.n_out; const size_t k = (size_t)std::max<int64_t>(nu, 1);
    Buf *b4[] = {&c->tb_name, &c->tb_piece, &c->tb_ttid, &c->tb_tstart, &c->tb_tend, &c->tb_etid, &c->tb_cov, &c->tb_ref, &c->tb_cnt};
    for (Buf *b : b4) NEED(*b, k * 4);
    NEED(c->tb_trev, k); NEED(c->tb_erev, k); NEED(c->tb_off, (k + 1) * 4);
    TabArgs a{};
    a.rows = *c->cur; a.ex = c->ex; a.list = c->novel; a.ref = c->u_ref.as<int32_t>(); a.rows_have_read_idx = c->have_batch ? 1 : 0;
    a.upd = merged_view(c->mg.o_cand, c->mg.o_cov, c->mg.o_tid, c->mg.o_start, c->mg.o_end, c->mg.o_fs, c->mg.o_le, nu);
    a.name_idx = c->tb_name.as<uint32_t>(); a.piece = c->tb_piece.as<int32_t>(); a.t_tid = c->tb_ttid.as<int32_t>(); a.t_start = c->tb_tstart.as<int32_t>();
    a.t_end = c->tb_tend.as<int32_t>(); a.e_tid = c->tb_etid.as<int32_t>(); a.cov = c->tb_cov.as<int32_t>(); a.ref_out = c->tb_ref.as<int32_t>();
    a.t_rev = c->tb_trev.as<uint8_t>(); a.e_rev = c->tb_erev.as<uint8_t>(); a.cnt = c->tb_cnt.as<uint32_t>(); a.off = c->tb_off.as<uint32_t>();
    int64_t ne = 0;
    if (nu) {
        if ((rc = ensure_tiles(c, nu))) return rc;
        tab_rows_kernel<<<(unsigned)((nu + 255) / 256), 256, 0, c->st>>>(a);
        launch_scan_sum_u32(a.cnt, c->tb_off.as<uint32_t>(), nu, c->tile_state.as<uint64_t>(), d_ticket(c), d_totals(c), c->st);
        CK(cudaGetLastError());
        uint64_t t; if ((rc = read_totals(c, &t, 1))) return rc;
        ne = (int64_t)t;
        NEED(c->tb_es, (size_t)std::max<int64_t>(ne, 1) * 4); NEED(c->tb_ee, (size_t)std::max<int64_t>(ne, 1) * 4);
        if (with_flags) NEED(c->tb_flag, (size_t)std::max<int64_t>(ne, 1));
        a.es = c->tb_es.as<int32_t>(); a.ee = c->tb_ee.as<int32_t>(); a.fl = with_flags ? c->tb_flag.as<uint8_t>() : nullptr;
        tab_exons_kernel<<<(unsigned)((nu + 255) / 256), 256, 0, c->st>>>(a);
        CK(cudaGetLastError());
    }
    *n_exon_out = ne;
    return LRB_OK;
}

extern "C" {

int lrb_update_fetch_table(lrb_ctx *c, lrb_trans_table *tab, lrb_bed_list *bed, int32_t *summary)
{
    if (!c) return LRB_E_ARG;
    if (!c->have_update) return fail(c, LRB_E_ARG, "lrb_update_fetch_table: update stage has not run");
    CK(cudaSetDevice(c->device));
    int rc;
    if (summary) memcpy(summary, c->summary, sizeof c->summary);
    if (bed) {
        if (!c->last_up.want_summary) return fail(c, LRB_E_ARG, "lrb_update_fetch_table: BED rows need want_summary");
        if ((rc = d2h(c, c->p[26], c->bd_tid.as<int32_t>(), (size_t)c->n_bed))) return rc;
        if ((rc = d2h(c, c->p[27], c->bd_s.as<int32_t>(), (size_t)c->n_bed))) return rc;
        if ((rc = d2h(c, c->p[28], c->bd_e.as<int32_t>(), (size_t)c->n_bed))) return rc;
        if ((rc = d2h(c, c->p[29], c->bd_sc.as<int32_t>(), (size_t)c->n_bed))) return rc;
        if ((rc = d2h(c, c->p[30], c->bd_ty.as<uint8_t>(), (size_t)c->n_bed))) return rc;
        if ((rc = d2h(c, c->p[31], c->bd_rv.as<uint8_t>(), (size_t)c->n_bed))) return rc;
        bed->n = c->n_bed; bed->tid = c->p[26].as<int32_t>(); bed->start = c->p[27].as<int32_t>(); bed->end = c->p[28].as<int32_t>();
        bed->score = c->p[29].as<int32_t>(); bed->type = c->p[30].as<uint8_t>(); bed->is_rev = c->p[31].as<uint8_t>();
    }
    if (tab) {
        int64_t ne = 0;
        if ((rc = lrbk::build_update_table(c, &ne, false))) return rc;
        const int64_t nu = c->mg.n_out; const size_t n = (size_t)nu;
        if ((rc = d2h(c, c->p[32], c->tb_name.as<uint32_t>(), n))) return rc;
        if ((rc = d2h(c, c->p[33], c->tb_piece.as<int32_t>(), n))) return rc;
        if ((rc = d2h(c, c->p[34], c->tb_ttid.as<int32_t>(), n))) return rc;
        if ((rc = d2h(c, c->p[35], c->tb_tstart.as<int32_t>(), n))) return rc;
        if ((rc = d2h(c, c->p[36], c->tb_tend.as<int32_t>(), n))) return rc;
        if ((rc = d2h(c, c->p[37], c->tb_trev.as<uint8_t>(), n))) return rc;
        if ((rc = d2h(c, c->p[38], c->tb_etid.as<int32_t>(), n))) return rc;
        if ((rc = d2h(c, c->p[39], c->tb_erev.as<uint8_t>(), n))) return rc;
        if ((rc = d2h(c, c->p[40], c->tb_cov.as<int32_t>(), n))) return rc;
        if ((rc = d2h(c, c->p[41], c->tb_ref.as<int32_t>(), n))) return rc;
        NEEDP(c->p[42], (n + 1) * 4);
        if (n) CK(cudaMemcpyAsync(c->p[42].p, c->tb_off.p, n * 4, cudaMemcpyDeviceToHost, c->st));
        if ((rc = d2h(c, c->p[43], c->tb_es.as<int32_t>(), (size_t)ne))) return rc;
        if ((rc = d2h(c, c->p[44], c->tb_ee.as<int32_t>(), (size_t)ne))) return rc;
        CK(cudaStreamSynchronize(c->st));
        c->p[42].as<uint32_t>()[n] = (uint32_t)ne;
        tab->n = nu; tab->name_idx = c->p[32].as<uint32_t>(); tab->piece = c->p[33].as<int32_t>(); tab->t_tid = c->p[34].as<int32_t>();
        tab->t_start = c->p[35].as<int32_t>(); tab->t_end = c->p[36].as<int32_t>(); tab->t_rev = c->p[37].as<uint8_t>(); tab->e_tid = c->p[38].as<int32_t>();
        tab->e_rev = c->p[39].as<uint8_t>(); tab->cov = c->p[40].as<int32_t>(); tab->ref_anno = c->p[41].as<int32_t>(); tab->exon_off = c->p[42].as<uint32_t>();
        tab->exon_start = c->p[43].as<int32_t>(); tab->exon_end = c->p[44].as<int32_t>();
    } else CK(cudaStreamSynchronize(c->st));
    return LRB_OK;
}

int lrb_unique_fetch(lrb_ctx *c, lrb_unique_result *out)
{
    if (!c || !out) return LRB_E_ARG;
    if (!c->have_unique) return fail(c, LRB_E_ARG, "lrb_unique_fetch: unique stage has not run");
    CK(cudaSetDevice(c->device));
    memset(out, 0, sizeof *out);
    int rc;
    if ((rc = fetch_exons(c, &out->ex, false, nullptr, 4))) return rc;
    if ((rc = fetch_merged(c, c->mg, &out->uniq, 19))) return rc;
    if ((rc = d2h(c, c->p[26], c->q_shared.as<uint32_t>(), (size_t)c->n_shared))) return rc;
    CK(cudaStreamSynchronize(c->st));
    out->n_shared = c->n_shared; out->shared_idx = c->p[26].as<uint32_t>();
    return LRB_OK;
}

// -------------------------------------------------------------------------------------------- one-call forms
int lrb_filter(lrb_ctx *c, const lrb_batch *b, const lrb_filter_params *p, lrb_filter_result *out)
{
    int rc;
    if ((rc = lrb_batch_upload(c, b))) return rc;
    if ((rc = lrb_filter_run(c, p))) return rc;
    return lrb_filter_fetch(c, out);
}
int lrb_bam2gtf(lrb_ctx *c, const lrb_batch *b, const lrb_exon_params *p, lrb_exon_result *out)
{
    int rc;
    if ((rc = lrb_batch_upload(c, b))) return rc;
    if ((rc = lrb_exon_run(c, p, 0))) return rc;
    return lrb_exon_fetch(c, out);
}
int lrb_update_gtf(lrb_ctx *c, const lrb_batch *b, const lrb_exon_params *ep, const lrb_update_params *up, lrb_update_result *out)
{
    int rc;
    if (b) { if ((rc = lrb_batch_upload(c, b))) return rc; if ((rc = lrb_exon_run(c, ep, 0))) return rc; }
    if ((rc = lrb_update_run(c, up))) return rc;
    return lrb_update_fetch(c, out);
}
int lrb_unique_gtf(lrb_ctx *c, const lrb_batch *b, const lrb_exon_params *ep, const lrb_update_params *up, lrb_unique_result *out)
{
    int rc;
    if (b) { if ((rc = lrb_batch_upload(c, b))) return rc; if ((rc = lrb_exon_run(c, ep, 0))) return rc; }
    if ((rc = lrb_unique_run(c, up))) return rc;
    return lrb_unique_fetch(c, out);
}

// ------------------------------------------------------------------------------------------------ measurement
int lrb_timing_enable(lrb_ctx *c, int on) { if (!c) return LRB_E_ARG; c->timing = on != 0; return LRB_OK; }
int lrb_timing_get(lrb_ctx *c, float ms[LRB_T_COUNT], int64_t *n_launches)
{
    if (!c) return LRB_E_ARG;
    if (ms) memcpy(ms, c->ms, sizeof c->ms);
    if (n_launches) *n_launches = c->launches_last;
    return LRB_OK;
}
int64_t lrb_launch_count(const lrb_ctx *c) { return c ? total_launches() - c->launches0 : 0; }
int lrb_mark(lrb_ctx *c, int slot)
{
    if (!c || slot < 0 || slot >= 8) return LRB_E_ARG;
    CK(cudaSetDevice(c->device));
    CK(cudaEventRecord(c->marks[slot], c->st));
    return LRB_OK;
}
int lrb_elapsed_ms(lrb_ctx *c, int a, int b, float *ms)
{
    if (!c || !ms || a < 0 || a >= 8 || b < 0 || b >= 8) return LRB_E_ARG;
    CK(cudaSetDevice(c->device));
    CK(cudaEventSynchronize(c->marks[b]));
    CK(cudaEventElapsedTime(ms, c->marks[a], c->marks[b]));
    return LRB_OK;
}
int lrb_update_diag(const lrb_ctx *c, int64_t *pieces, int64_t *replays)
{
    if (!c) return LRB_E_ARG;
    if (pieces) *pieces = c->n_xlocus_pieces;
    if (replays) *replays = c->n_xlocus_replays;
    return LRB_OK;
}
void *lrb_host_alloc(size_t bytes) { void *p = nullptr; if (cudaMallocHost(&p, bytes ? bytes : 1) != cudaSuccess) { cudaGetLastError(); return nullptr; } return p; }
void lrb_host_free(void *p) { if (p) cudaFreeHost(p); }

// ----------------------------------------------------------------------------------------------------- shards
int lrb_shard_cuts_weighted(const int32_t *tid, const int32_t *start, const int32_t *end, const int64_t *weight, int64_t n, int n_shards, int64_t *cuts)
{
    if (n_shards <= 0 || !cuts || n < 0) return LRB_E_ARG;
    // a cut is exact where the read starts beyond every earlier end on its chromosome;
    // take the first such position at or after each ideal boundary (k/n_shards of the total weight)
    cuts[0] = 0; cuts[n_shards] = n;
    __int128 total = 0;
    if (weight) for (int64_t i = 0; i < n; ++i) total += weight[i]; else total = n;
    uint64_t run = 0; int k = 1; __int128 acc = 0, want = n_shards > 1 ? total / n_shards : total;
    for (int64_t i = 0; i < n && k < n_shards; ++i) {
        uint64_t ks = ((uint64_t)(uint32_t)(tid[i] + 1) << 32) | (uint32_t)start[i];
        if (acc >= want && i > cuts[k - 1] && ks > run) { cuts[k++] = i; want = total * k / n_shards; }
        uint64_t ke = ((uint64_t)(uint32_t)(tid[i] + 1) << 32) | (uint32_t)end[i];
        if (ke > run) run = ke;
        acc += weight ? weight[i] : 1;
    }
    for (; k < n_shards; ++k) cuts[k] = n;
    return LRB_OK;
}
int lrb_shard_cuts(const int32_t *tid, const int32_t *start, const int32_t *end, int64_t n, int n_shards, int64_t *cuts)
{
    return lrb_shard_cuts_weighted(tid, start, end, nullptr, n, n_shards, cuts);
}

}  // extern "C"
