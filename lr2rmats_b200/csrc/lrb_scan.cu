// lrb_scan.cu -- K1: the fused CIGAR pass (filter statistics + exon chains) and the qname-run selection.
//
// Replaces, for a whole batch at once:
//   gtf_filter()      bam_filter.c:61-86   (coverage / NM identity / remove-GTF overlap, score)
//   remove_overlap()  bam_filter.c:49-59
//   gen_exon()        bam2gtf.c:31-78      (CIGAR walk -> exon chain)
//   the best/second-best qname-run state machine of bam_filter(), bam_filter.c:129-159
//
// Layout: one CTA owns a tile of `reads_per_tile` consecutive records.  Their CIGAR words are one contiguous range of
// the pool, staged into shared memory with coalesced loads and walked from there -- by one thread per read when
// CIGARs are short (Iso-Seq like), or by one warp per read with ballot / shuffle prefix scans when they are long
// (ONT like).  Exons are produced in a second walk over the staged words once the tile's output offset is known from
// a decoupled look-back over (rows, exons), staged in shared memory and written out coalesced: the CIGAR pool is read
// from HBM exactly once and every output array is written exactly once, in read order.
#include "lrb_common.cuh"
#include "lrb_kernels.cuh"

namespace lrbk {

static int64_t g_launches = 0;
int64_t count_launches() { return g_launches; }
#define LRB_COUNT_LAUNCH() (++g_launches)

static constexpr int SCAN_THREADS = 256;

enum { OP_M = 0, OP_I, OP_D, OP_N, OP_S, OP_H, OP_P, OP_EQ, OP_X, OP_B };

LRB_DEVINL bool op_ref(unsigned op) { return (0x18Du >> op) & 1u; }   // M D N = X consume reference (bits 0,2,3,7,8)

struct WalkStats { int n_exon, intron_n, del_len, ref_len, first_start, last_end; };

// ---- sequential walk (one thread).  EMIT writes exons to es/ee.
template <bool EMIT>
LRB_DEVINL void walk_seq(const uint32_t *c, int n_c, int pos, const lrb_exon_params &ep, int *es, int *ee, WalkStats &w)
{
    int n = 0, start = pos + 1, end = pos, intron = 0, del = 0, first_start = pos + 1;
    for (int i = 0; i < n_c; ++i) {
        uint32_t x = c[i]; int l = (int)(x >> 4); unsigned op = x & 15u;
        bool cut = (op == OP_N && l >= ep.min_intron) || (op == OP_D && l > ep.max_delet);
        if (op == OP_N) ++intron; else if (op == OP_D) del += l;
        if (cut) {
            if (n == 0 || (end - start + 1) >= ep.min_exon) { if (EMIT) { es[n] = start; ee[n] = end; } ++n; }
            start = end + l + 1;
        }
        if (op_ref(op)) end += l;
    }
    if (EMIT) { es[n] = start; ee[n] = end; }
    ++n;
    w.n_exon = n; w.intron_n = intron; w.del_len = del; w.ref_len = end - pos; w.first_start = first_start; w.last_end = end;
}

// ---- sequential walk that also keeps the exons: exon k (all but the last) is parked in the words 2k, 2k+1 of the read's own
// staged CIGAR, which the walk has consumed by then whenever cuts are not adjacent and the CIGAR does not start with one
// (word 2k+1 <= index of the cut that closes exon k).  A CIGAR that breaks this sets *ovf and is walked again later.
LRB_DEVINL void walk_seq_inplace(uint32_t *c, int n_c, int pos, const lrb_exon_params &ep, WalkStats &w, bool *ovf, int *last_start)
{
    int n = 0, start = pos + 1, end = pos, intron = 0, del = 0; bool over = false;
    for (int i = 0; i < n_c; ++i) {
        uint32_t x = c[i]; int l = (int)(x >> 4); unsigned op = x & 15u;
        bool cut = (op == OP_N && l >= ep.min_intron) || (op == OP_D && l > ep.max_delet);
        if (op == OP_N) ++intron; else if (op == OP_D) del += l;
        if (cut) {
            if (n == 0 || (end - start + 1) >= ep.min_exon) {
                if (2 * n + 1 <= i && !over) { c[2 * n] = (uint32_t)start; c[2 * n + 1] = (uint32_t)end; } else over = true;
                ++n;
            }
            start = end + l + 1;
        }
        if (op_ref(op)) end += l;
    }
    ++n;
    *ovf = over; *last_start = start;
    w.n_exon = n; w.intron_n = intron; w.del_len = del; w.ref_len = end - pos; w.first_start = pos + 1; w.last_end = end;
}

// ---- cooperative walk (one warp, lanes over ops, 32 ops per step)
template <bool EMIT>
LRB_DEVINL void walk_warp(const uint32_t *c, int n_c, int pos, const lrb_exon_params &ep, int *es, int *ee, WalkStats &w)
{
    const int lane = lane_id();
    int carry_ref = 0;               // reference bases consumed by earlier steps
    int last_cut_after = pos;        // `end` right after the latest cut op (start of the open exon - 1)
    bool seen_cut = false;
    int n = 0, intron = 0, del = 0;
    for (int base = 0; base < n_c; base += 32) {
        int i = base + lane;
        uint32_t x = i < n_c ? c[i] : 0xFu;                     // op 15: consumes nothing
        int l = (int)(x >> 4); unsigned op = x & 15u;
        int rc = (op <= 8 && op_ref(op)) ? l : 0;
        bool cut = (op == OP_N && l >= ep.min_intron) || (op == OP_D && l > ep.max_delet);
        intron += __popc(__ballot_sync(FULL, op == OP_N));
        int dl = op == OP_D ? l : 0;
        int inc = rc;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { int t = __shfl_up_sync(FULL, inc, o); int d = __shfl_up_sync(FULL, dl, o); if (lane >= o) { inc += t; dl += d; } }
        del += __shfl_sync(FULL, dl, 31);
        int end_after = pos + carry_ref + inc, end_before = end_after - rc;
        unsigned cutmask = __ballot_sync(FULL, cut);
        unsigned prev = cutmask & ((1u << lane) - 1u);
        int src = prev ? 31 - __clz(prev) : 0;
        int prev_after = __shfl_sync(FULL, end_after, src);
        int start = (prev ? prev_after : last_cut_after) + 1;
        bool first_cut = cut && !seen_cut && prev == 0;
        bool emit = cut && (first_cut || (end_before - start + 1) >= ep.min_exon);
        unsigned emask = __ballot_sync(FULL, emit);
        if (EMIT && emit) { int k = n + __popc(emask & ((1u << lane) - 1u)); es[k] = start; ee[k] = end_before; }
        n += __popc(emask);
        if (cutmask) { last_cut_after = __shfl_sync(FULL, end_after, 31 - __clz(cutmask)); seen_cut = true; }
        carry_ref += __shfl_sync(FULL, inc, 31);
    }
    if (EMIT && lane == 0) { es[n] = last_cut_after + 1; ee[n] = pos + carry_ref; }
    ++n;
    w.n_exon = n; w.intron_n = intron; w.del_len = del; w.ref_len = carry_ref; w.first_start = pos + 1; w.last_end = pos + carry_ref;
}

// remove_overlap() through the per-tid index (see DRmIndex)
LRB_DEVINL bool rm_hit(const DRmIndex &rm, int tid, int pos, int rlen)
{
    if (rm.n_groups == 0) return false;
    int lo = 0, hi = rm.n_groups;
    while (lo < hi) { int m = (lo + hi) >> 1; if (rm.g_tid[m] < tid) lo = m + 1; else hi = m; }
    if (lo >= rm.n_groups || rm.g_tid[lo] != tid) return false;
    int b = rm.g_off[lo], e = rm.g_off[lo + 1], qe = pos + rlen - 1;
    // last entry with start <= qe
    int l2 = b, h2 = e;
    while (l2 < h2) { int m = (l2 + h2) >> 1; if (rm.start[m] <= qe) l2 = m + 1; else h2 = m; }
    if (l2 == b) return false;
    return rm.pmax_end[l2 - 1] >= pos;            // !(pos > end): 0-based pos against 1-based end, as the reference
}

// gtf_filter() predicate after the walk (bam_filter.c:73-84); mixed float/double compares kept as in C
LRB_DEVINL bool filter_pass(const ScanArgs &a, int64_t r, uint32_t c0, uint32_t c1, int n_c, const WalkStats &w, int *score)
{
    if (a.b.flag[r] & 4) return false;
    int l_qseq = a.b.l_qseq[r], qlen = l_qseq;
    if (n_c > 0) {
        unsigned op0 = c0 & 15u, op1 = c1 & 15u;
        if (op0 == OP_S || op0 == OP_H) qlen -= (int)(c0 >> 4);
        if (n_c > 1 && (op1 == OP_S || op1 == OP_H)) qlen -= (int)(c1 >> 4);
    }
    if (__ddiv_rn((double)qlen, (double)l_qseq) < (double)a.fp.cov_rate) return false;
    int sc = qlen - a.b.nm[r] + w.del_len;
    if ((float)sc < __fmul_rn(a.fp.map_qual, (float)qlen)) return false;
    if (rm_hit(a.rm, a.b.tid[r], a.b.pos[r], w.ref_len)) return false;
    *score = sc;
    return true;
}

// Short CIGARs (Iso-Seq like): one thread per read on a staged tile.  Long CIGARs go through cigar_stream_kernel below.
__global__ void __launch_bounds__(SCAN_THREADS) cigar_scan_kernel(ScanArgs a)
{
    extern __shared__ __align__(16) uint32_t smem[];
    uint32_t *s_words = smem;                                   // stage_words
    __shared__ uint32_t s_scan[33];
    __shared__ int s_cnt[SCAN_THREADS], s_start[SCAN_THREADS], s_end[SCAN_THREADS];
    __shared__ uint8_t s_mask[SCAN_THREADS];
    __shared__ uint32_t s_tile;
    __shared__ uint64_t s_excl;

    const int tid = threadIdx.x, R = a.reads_per_tile;
    if (tid == 0) s_tile = atomicAdd(a.ticket, 1u);
    __syncthreads();
    const int tile = (int)s_tile;
    const int64_t r0 = (int64_t)tile * R, r1 = min(a.b.n, r0 + R);
    const int nr = (int)(r1 - r0);
    const uint64_t w_lo = a.b.cigar_off[r0], w_hi = a.b.cigar_off[r1];
    const uint64_t nw = w_hi - w_lo;
    const bool staged = nw <= (uint64_t)a.stage_words;
    const bool do_filter = a.mode != 1, do_exon = a.mode != 0;
    if (staged) {
        // coalesced stage: scalar head up to 16-byte alignment, then 128-bit streaming loads
        const uint32_t *src = a.b.cigar + w_lo;
        uint32_t head = (uint32_t)((4 - (w_lo & 3)) & 3); if (head > nw) head = (uint32_t)nw;
        if ((uint32_t)tid < head) s_words[tid] = ldg_stream_u32(src + tid);
        uint32_t nv = ((uint32_t)nw - head) >> 2;
        const uint4 *v = (const uint4 *)(src + head);
        for (uint32_t i = tid; i < nv; i += SCAN_THREADS) {
            uint4 q = ldg_stream_u4(v + i);
            uint32_t o = head + 4 * i;
            s_words[o] = q.x; s_words[o + 1] = q.y; s_words[o + 2] = q.z; s_words[o + 3] = q.w;
        }
        uint32_t tail = head + 4 * nv;
        if (tail + tid < nw) s_words[tail + tid] = ldg_stream_u32(src + tail + tid);
    }
    s_cnt[tid] = 0; s_mask[tid] = 0;
    __syncthreads();

    // ---- walk 1: statistics, exon count, filter predicate
    auto read_ptr = [&](int64_t r) -> const uint32_t * {
        const uint64_t off = a.b.cigar_off[r];
        return staged ? (s_words + (off - w_lo)) : (a.b.cigar + off);
    };
    bool parked = false, park_ovf = false; int park_last_start = 0;      // exons parked in the staged words by walk 1
    if (tid < nr) {
        int64_t r = r0 + tid;
        const uint32_t *c = read_ptr(r); int n_c = (int)(a.b.cigar_off[r + 1] - a.b.cigar_off[r]);
        const uint32_t c0 = n_c > 0 ? c[0] : 0u, c1 = n_c > 0 ? c[n_c - 1] : 0u;
        WalkStats w;
        if (staged && do_exon) { parked = true; walk_seq_inplace(const_cast<uint32_t *>(c), n_c, a.b.pos[r], a.ep, w, &park_ovf, &park_last_start); }
        else walk_seq<false>(c, n_c, a.b.pos[r], a.ep, nullptr, nullptr, w);
        bool mask;
        if (do_filter) {
            int sc = 0; bool p = filter_pass(a, r, c0, c1, n_c, w, &sc);
            a.pass[r] = p ? 1 : 0;
            if (p) { a.score[r] = sc; a.intron_n[r] = w.intron_n; }
            mask = p;
        } else mask = a.sel_mask ? (a.sel_mask[r] != 0) : true;
        const bool unmapped = (a.b.flag[r] & 4) != 0;
        s_mask[tid] = mask ? 1 : 0;
        s_cnt[tid] = (mask && do_exon && !unmapped) ? w.n_exon : 0;
        s_start[tid] = w.first_start; s_end[tid] = w.last_end;
    }
    __syncthreads();

    // ---- tile offsets: block scan of (rows, exons) + look-back across tiles
    uint32_t my_row = s_mask[tid], my_ex = (uint32_t)s_cnt[tid], rows_total, ex_total;
    uint32_t row_excl = block_excl_sum(my_row, s_scan, &rows_total);
    uint32_t ex_excl = block_excl_sum(my_ex, s_scan, &ex_total);
    if (warp_id() == 0) {
        uint64_t e = lookback_exclusive(a.tile_state, tile, pack_pair(rows_total, ex_total), OpAdd());
        if (lane_id() == 0) s_excl = e;
    }
    __syncthreads();
    const uint32_t row_base = pair_hi(s_excl), ex_base = pair_lo(s_excl);
    if (r1 == a.b.n && tid == 0) { a.totals[0] = (uint64_t)row_base + rows_total; a.totals[1] = (uint64_t)ex_base + ex_total; }

    // ---- row records
    if (tid < nr && my_row) {
        int64_t r = r0 + tid; uint32_t row = a.rows_by_record ? (uint32_t)r : row_base + row_excl;
        if ((int64_t)row < a.rows.cap) {
            a.rows.read_idx[row] = (uint32_t)r;
            if (do_exon) {
                int8_t xs = a.b.xs[r];
                a.rows.tid[row] = a.b.tid[r];
                a.rows.is_rev[row] = xs == 0 ? ((a.b.flag[r] & 16) != 0) : (xs == '+' ? 0 : 1);    // bam2gtf.c:35-37
                a.rows.start[row] = s_start[tid]; a.rows.end[row] = s_end[tid];
                a.rows.ex_beg[row] = ex_base + ex_excl; a.rows.ex_n[row] = my_ex;
            }
        }
    }
    if (!do_exon || ex_total == 0) return;

    // ---- walk 2: each read's parked exons move straight to the pools (no staging buffer: 6 resident CTAs per SM cover
    // the look-back waits)
    const bool room = (int64_t)ex_base + ex_total <= a.ex.cap;   // host re-runs with a larger pool otherwise
    if (!room) return;
    if (tid < nr && my_ex) {
        int64_t r = r0 + tid;
        const uint32_t *c = read_ptr(r); int n_c = (int)(a.b.cigar_off[r + 1] - a.b.cigar_off[r]);
        int *es = a.ex.es + ex_base + ex_excl, *ee = a.ex.ee + ex_base + ex_excl;
        if (parked && !park_ovf) {                            // move the parked exons; the open one is still in registers
            const int last = (int)my_ex - 1;
            for (int k = 0; k < last; ++k) { es[k] = (int)c[2 * k]; ee[k] = (int)c[2 * k + 1]; }
            es[last] = park_last_start; ee[last] = s_end[tid];
        } else {
            if (parked) c = a.b.cigar + a.b.cigar_off[r];     // the staged copy is partly overwritten: walk the pool
            WalkStats w; walk_seq<true>(c, n_c, a.b.pos[r], a.ep, es, ee, w);
        }
    }
}

// ------------------------------------------------------------------------------------ long CIGARs: streaming scan
// ONT-like reads carry hundreds of ops, so the pass is a stream over the CIGAR pool with a segmented scan on top.
//
//   * a CTA owns a tile of <= ST_R consecutive reads = ONE contiguous range of the pool; its 8 warps split that range at
//     read boundaries into spans of about equal op counts, so a read never straddles two warps and no block-wide
//     synchronisation happens while the ops stream;
//   * every warp pulls its span through its own shared-memory ring with 1-D TMA bulk copies (cp.async.bulk + mbarrier
//     complete_tx, ST_STAGES stages in flight: the loads of the next rounds run under the scan of the current one) and
//     consumes it in rounds of 256 ops, 8 consecutive ops per lane (two conflict-free 128-bit shared loads);
//   * what a sequential walk carries from op to op is a flat prefix over the span, read boundaries ignored: reference
//     bases consumed (32-bit, modular), deletion bases, N ops, cut ops -- two warp scans per round; per-read values are
//     differences of that prefix between the read's last and first op, so the pool is read exactly ONCE;
//   * the flat prefix of the reference bases is written back over the ops in the ring, so the (few) cut ops find the
//     exon end in front of them and the exon start behind them with two shared loads; exon k of a read lands in the warp's
//     staging slot (cuts so far in the span) + (reads so far in the span), whatever the read boundaries are;
//   * the tile's reads are then finished by one thread each (filter predicate, short internal exons dropped as
//     bam2gtf.c:45 does), placed by the same (rows, exons) look-back as the short-CIGAR kernel and copied to the pools.
// A tile whose exons overflow the staging slots is redone by walk_warp (one warp per read, two walks).
static constexpr int ST_THREADS = 256, ST_WARPS = ST_THREADS / 32;
static constexpr int ST_R = 128;                    // reads per tile (upper bound)
static constexpr int ST_ROUND = 256;                // ops per warp round = 8 per lane
static constexpr int ST_CHUNK = 256;                // words per ring stage (one round)
static constexpr int ST_STAGES = 4;
static constexpr int ST_EXW = 256;                  // exon staging slots per warp

LRB_DEVINL uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
LRB_DEVINL void mbar_init(uint64_t *bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory"); }
LRB_DEVINL void mbar_expect_tx(uint64_t *bar, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory"); }
LRB_DEVINL void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
LRB_DEVINL bool mbar_try_wait(uint64_t *bar, uint32_t parity)
{
    uint32_t ok;
    asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}" : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}

struct StreamSmem {
    alignas(128) uint32_t ring[ST_WARPS][ST_STAGES][ST_CHUNK];
    int es[ST_WARPS][ST_EXW], ee[ST_WARPS][ST_EXW];
    alignas(8) uint64_t bar[ST_WARPS][ST_STAGES];
    int off[ST_R + 1], pos[ST_R];
    uint2 dslot[ST_R];                              // x: D = pos - (flat reference prefix at the read's first op), y: reads before it in the span
    int ref_len[ST_R], del_len[ST_R], intron_n[ST_R];
    uint16_t ncut[ST_R], sbeg[ST_R]; uint8_t wof[ST_R];
    int cnt[ST_THREADS], start[ST_THREADS], end[ST_THREADS];
    uint8_t mask[ST_THREADS];
    uint32_t scan[33]; uint32_t tile; int ovf; uint64_t excl;
};

__global__ void __launch_bounds__(ST_THREADS, 3) cigar_stream_kernel(ScanArgs a)
{
    extern __shared__ __align__(128) uint8_t st_raw[];
    StreamSmem &S = *reinterpret_cast<StreamSmem *>(st_raw);
    const int tid = threadIdx.x, lane = lane_id(), w = warp_id(), R = a.reads_per_tile;
    if (tid == 0) { S.tile = atomicAdd(a.ticket, 1u); S.ovf = 0; }
    if (lane == 0) for (int s = 0; s < ST_STAGES; ++s) mbar_init(&S.bar[w][s], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncthreads();
    const int tile = (int)S.tile;
    const int64_t r0 = (int64_t)tile * R, r1 = min(a.b.n, r0 + R);
    const int nr = (int)(r1 - r0);
    const uint64_t w_lo = a.b.cigar_off[r0], w_hi = a.b.cigar_off[r1];
    const bool do_filter = a.mode != 1, do_exon = a.mode != 0;
    // a tile of more than 2^30 ops (128 reads of 65535 ops are 8.4 M) cannot exist with 16-bit n_cigar; guard anyway
    const bool huge = w_hi - w_lo >= (1ull << 30);
    const int nw = huge ? 0 : (int)(w_hi - w_lo);
    if (tid <= nr) S.off[tid] = huge ? 0 : (int)(a.b.cigar_off[r0 + tid] - w_lo);
    if (tid < nr) S.pos[tid] = a.b.pos[r0 + tid];
    S.cnt[tid] = 0; S.mask[tid] = 0;
    if (huge && tid == 0) S.ovf = 1;
    __syncthreads();

    // ---- phase A: every warp streams its span of reads
    if (nw > 0) {
        // span of warp w: reads [rs, re) = those that START in its share of the op range
        auto first_read_at = [&](int t) { int lo = 0, hi = nr; while (lo < hi) { const int m = (lo + hi) >> 1; if (S.off[m] < t) lo = m + 1; else hi = m; } return lo; };
        const int rs = w == 0 ? 0 : first_read_at((int)((int64_t)nw * w / ST_WARPS));
        const int re = w == ST_WARPS - 1 ? nr : first_read_at((int)((int64_t)nw * (w + 1) / ST_WARPS));
        const int a_op = S.off[rs], b_op = S.off[re];                    // ops [a_op, b_op) of the tile
        if (a_op < b_op) {
            uint32_t *ring = &S.ring[w][0][0]; uint64_t *bar = &S.bar[w][0];
            int *xes = S.es[w], *xee = S.ee[w];
            const uint64_t g_abs = (w_lo + (uint64_t)a_op) & ~3ull;      // 16-byte aligned start of the stream
            const int lead = (int)(w_lo + (uint64_t)a_op - g_abs);       // pad words in front of the span (0..3)
            const int n_words = lead + (b_op - a_op);                    // words of the stream
            const int n_chunks = (n_words + ST_CHUNK - 1) / ST_CHUNK;
            const uint32_t *src = a.b.cigar + g_abs;
            auto issue = [&](int c) {
                const int s = c % ST_STAGES;
                int words = n_words - c * ST_CHUNK; words = words > ST_CHUNK ? ST_CHUNK : ((words + 3) & ~3);
                mbar_expect_tx(&bar[s], (uint32_t)words * 4u);
                bulk_g2s(ring + s * ST_CHUNK, src + (size_t)c * ST_CHUNK, (uint32_t)words * 4u, &bar[s]);
            };
            if (lane == 0) for (int c = 0; c < ST_STAGES && c < n_chunks; ++c) issue(c);
            // warp-uniform state of the read being walked
            int r = rs; while (S.off[r + 1] == a_op) ++r;                // first read with ops (a_op < b_op: it exists)
            int nb = S.off[r + 1];                                       // its end
            uint32_t baseS = 0, baseD = 0, baseN = 0, baseC = 0;         // flat prefixes at its first op
            int ridx = 0;                                                // reads with ops started before it in the span
            uint32_t carryS = 0, carryD = 0, carryN = 0, carryC = 0;     // flat prefixes at the start of the round
            if (lane == 0) {
                S.dslot[r] = make_uint2((uint32_t)S.pos[r], 0u); S.sbeg[r] = 0; S.wof[r] = (uint8_t)w;
                if (do_exon) xes[0] = S.pos[r] + 1;
            }
            // per-lane read tracking for the cut ops
            int rL = r, nbL = nb;
            bool done = false;
            for (int c = 0; c < n_chunks && !done; ++c) {
                const int s = c % ST_STAGES;
                while (!mbar_try_wait(&bar[s], (uint32_t)(c / ST_STAGES) & 1u)) { }
                uint32_t *rw = ring + s * ST_CHUNK;
                const int g = a_op - lead + c * ST_CHUNK;                // tile op index of the round's first word
                const int i0 = g + 8 * lane;
                uint4 q0 = *reinterpret_cast<const uint4 *>(rw + 8 * lane), q1 = *reinterpret_cast<const uint4 *>(rw + 8 * lane + 4);
                uint32_t x[8] = {q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, q1.z, q1.w};
                uint32_t sp[8], dp[8], cm = 0, nm = 0, sacc = 0, dacc = 0;
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    const int i = i0 + k;
                    const uint32_t xv = (i >= a_op && i < b_op) ? x[k] : 0xFu;   // outside the span: op 15 consumes nothing
                    const uint32_t l = xv >> 4, op = xv & 15u;
                    if ((0x18Du >> op) & 1u) sacc += l;
                    if (op == OP_D) dacc += l;
                    const bool isn = op == OP_N;
                    const bool cut = (isn && (int)l >= a.ep.min_intron) || (op == OP_D && (int)l > a.ep.max_delet);
                    nm |= (isn ? 1u : 0u) << k; cm |= (cut ? 1u : 0u) << k;
                    sp[k] = sacc; dp[k] = dacc;
                }
                // two warp scans: reference bases; (deletion bases : N ops : cut ops) packed
                uint32_t incS = sacc;
                unsigned long long bpk = ((unsigned long long)dacc << 32) | ((unsigned long long)__popc(nm) << 16) | (unsigned long long)__popc(cm);
                unsigned long long incB = bpk;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const uint32_t t = __shfl_up_sync(FULL, incS, o); const unsigned long long u = __shfl_up_sync(FULL, incB, o);
                    if (lane >= o) { incS += t; incB += u; }
                }
                const uint32_t exS = carryS + incS - sacc;
                const unsigned long long exB = incB - bpk;
                const uint32_t exD = carryD + (uint32_t)(exB >> 32), exN = carryN + (uint32_t)((exB >> 16) & 0xffffu), exC = carryC + (uint32_t)(exB & 0xffffu);
                // the flat reference prefix BEHIND every op goes back over the ops (the cut ops read it from there)
#pragma unroll
                for (int k = 0; k < 8; ++k) sp[k] += exS;
                *reinterpret_cast<uint4 *>(rw + 8 * lane) = make_uint4(sp[0], sp[1], sp[2], sp[3]);
                *reinterpret_cast<uint4 *>(rw + 8 * lane + 4) = make_uint4(sp[4], sp[5], sp[6], sp[7]);
                // ---- read boundaries of this round, warp-uniform: position nb lies behind op nb - 1
                const int gend = g + ST_ROUND;
                while (nb <= gend) {
                    const int p = nb - 1 - g, ol = p >> 3, ok = p & 7;   // owner lane / slot of the read's last op
                    uint32_t vS = sp[0], vD = dp[0];
                    if (ok == 1) { vS = sp[1]; vD = dp[1]; } else if (ok == 2) { vS = sp[2]; vD = dp[2]; } else if (ok == 3) { vS = sp[3]; vD = dp[3]; }
                    else if (ok == 4) { vS = sp[4]; vD = dp[4]; } else if (ok == 5) { vS = sp[5]; vD = dp[5]; } else if (ok == 6) { vS = sp[6]; vD = dp[6]; }
                    else if (ok == 7) { vS = sp[7]; vD = dp[7]; }
                    const uint32_t low = (2u << ok) - 1u;
                    const uint32_t Sb = __shfl_sync(FULL, vS, ol), Db = __shfl_sync(FULL, exD + vD, ol);
                    const uint32_t Nb = __shfl_sync(FULL, exN + __popc(nm & low), ol), Cb = __shfl_sync(FULL, exC + __popc(cm & low), ol);
                    if (lane == 0) {
                        S.ref_len[r] = (int)(Sb - baseS); S.del_len[r] = (int)(Db - baseD); S.intron_n[r] = (int)(Nb - baseN);
                        const uint32_t nc = Cb - baseC; S.ncut[r] = (uint16_t)(nc > 0xffffu ? 0xffffu : nc);
                        if (do_exon) { const uint32_t sl = Cb + (uint32_t)ridx; if (sl < (uint32_t)ST_EXW) xee[sl] = S.pos[r] + (int)(Sb - baseS); else S.ovf = 1; }
                    }
                    int rn = r + 1; while (rn < re && S.off[rn + 1] == nb) ++rn;          // next read with ops
                    if (rn >= re) { done = true; break; }
                    r = rn; nb = S.off[r + 1]; baseS = Sb; baseD = Db; baseN = Nb; baseC = Cb; ++ridx;
                    if (lane == 0) {
                        const uint32_t sl = Cb + (uint32_t)ridx;
                        S.dslot[r] = make_uint2((uint32_t)S.pos[r] - Sb, (uint32_t)ridx); S.sbeg[r] = (uint16_t)(sl < (uint32_t)ST_EXW ? sl : ST_EXW); S.wof[r] = (uint8_t)w;
                        if (do_exon) { if (sl < (uint32_t)ST_EXW) xes[sl] = S.pos[r] + 1; else S.ovf = 1; }
                    }
                }
                __syncwarp();
                // ---- cut ops: exon end in front of the cut, exon start behind it
                if (do_exon) {
                    uint32_t m = cm;
                    while (m) {
                        const int k = __ffs(m) - 1; m &= m - 1;
                        const int i = i0 + k, wi = 8 * lane + k;
                        while (i >= nbL) { ++rL; nbL = S.off[rL + 1]; }
                        const uint2 ds = S.dslot[rL];
                        const uint32_t s_after = rw[wi], s_before = wi ? rw[wi - 1] : carryS;
                        const uint32_t sl = exC + __popc(cm & ((1u << k) - 1u)) + ds.y;
                        if (sl + 1 < (uint32_t)ST_EXW) { xee[sl] = (int)(ds.x + s_before); xes[sl + 1] = (int)(ds.x + s_after) + 1; }
                        else S.ovf = 1;
                    }
                }
                // ---- carries into the next round
                const unsigned long long totB = __shfl_sync(FULL, incB, 31);
                carryS += __shfl_sync(FULL, incS, 31); carryD += (uint32_t)(totB >> 32); carryN += (uint32_t)((totB >> 16) & 0xffffu); carryC += (uint32_t)(totB & 0xffffu);
                __syncwarp();
                if (lane == 0 && c + ST_STAGES < n_chunks) { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); issue(c + ST_STAGES); }
            }
        }
    }
    __syncthreads();
    const bool ovf = S.ovf != 0;

    // ---- phase B: one thread finishes one read
    auto finish_read = [&](int li, int64_t rr, uint32_t c0, uint32_t c1, int n_c, const WalkStats &ws) {
        bool mask;
        if (do_filter) {
            int sc = 0; bool p = filter_pass(a, rr, c0, c1, n_c, ws, &sc);
            a.pass[rr] = p ? 1 : 0;
            if (p) { a.score[rr] = sc; a.intron_n[rr] = ws.intron_n; }
            mask = p;
        } else mask = a.sel_mask ? (a.sel_mask[rr] != 0) : true;
        const bool unmapped = (a.b.flag[rr] & 4) != 0;
        S.mask[li] = mask ? 1 : 0;
        S.cnt[li] = (mask && do_exon && !unmapped) ? ws.n_exon : 0;
        S.start[li] = ws.first_start; S.end[li] = ws.last_end;
    };
    if (!ovf) {
        if (tid < nr) {
            const int64_t rr = r0 + tid;
            const int n_c = S.off[tid + 1] - S.off[tid];
            const uint32_t *cg = a.b.cigar + w_lo;
            const uint32_t c0 = n_c > 0 ? cg[S.off[tid]] : 0u, c1 = n_c > 0 ? cg[S.off[tid + 1] - 1] : 0u;
            WalkStats ws;
            ws.first_start = S.pos[tid] + 1;
            if (n_c > 0) {
                ws.ref_len = S.ref_len[tid]; ws.del_len = S.del_len[tid]; ws.intron_n = S.intron_n[tid];
                int n = (int)S.ncut[tid] + 1;
                if (do_exon && n > 2) {                          // short internal exons vanish (bam2gtf.c:45), the first and the last stay
                    int *xs = S.es[S.wof[tid]] + S.sbeg[tid], *xe = S.ee[S.wof[tid]] + S.sbeg[tid];
                    int o = 1;
                    for (int k = 1; k < n - 1; ++k) if (xe[k] - xs[k] + 1 >= a.ep.min_exon) { xs[o] = xs[k]; xe[o] = xe[k]; ++o; }
                    xs[o] = xs[n - 1]; xe[o] = xe[n - 1]; n = o + 1;
                }
                ws.n_exon = n;
            } else { ws.ref_len = 0; ws.del_len = 0; ws.intron_n = 0; ws.n_exon = 1; }
            ws.last_end = S.pos[tid] + ws.ref_len;
            finish_read(tid, rr, c0, c1, n_c, ws);
        }
    } else {
        for (int li = w; li < nr; li += ST_WARPS) {
            const int64_t rr = r0 + li;
            const uint32_t *c = a.b.cigar + a.b.cigar_off[rr]; const int n_c = (int)(a.b.cigar_off[rr + 1] - a.b.cigar_off[rr]);
            WalkStats ws; walk_warp<false>(c, n_c, a.b.pos[rr], a.ep, nullptr, nullptr, ws);
            if (lane == 0) finish_read(li, rr, n_c > 0 ? c[0] : 0u, n_c > 0 ? c[n_c - 1] : 0u, n_c, ws);
        }
    }
    __syncthreads();

    // ---- tile offsets: block scan of (rows, exons) + look-back across tiles
    const uint32_t my_row = S.mask[tid], my_ex = (uint32_t)S.cnt[tid];
    uint32_t rows_total, ex_total;
    const uint32_t row_excl = block_excl_sum(my_row, S.scan, &rows_total);
    const uint32_t ex_excl = block_excl_sum(my_ex, S.scan, &ex_total);
    if (w == 0) {
        const uint64_t e = lookback_exclusive(a.tile_state, tile, pack_pair(rows_total, ex_total), OpAdd());
        if (lane == 0) S.excl = e;
    }
    __syncthreads();
    const uint32_t row_base = pair_hi(S.excl), ex_base = pair_lo(S.excl);
    if (r1 == a.b.n && tid == 0) { a.totals[0] = (uint64_t)row_base + rows_total; a.totals[1] = (uint64_t)ex_base + ex_total; }
    if (tid < nr && my_row) {
        const int64_t rr = r0 + tid; const uint32_t row = a.rows_by_record ? (uint32_t)rr : row_base + row_excl;
        if ((int64_t)row < a.rows.cap) {
            a.rows.read_idx[row] = (uint32_t)rr;
            if (do_exon) {
                const int8_t xs = a.b.xs[rr];
                a.rows.tid[row] = a.b.tid[rr];
                a.rows.is_rev[row] = xs == 0 ? ((a.b.flag[rr] & 16) != 0) : (xs == '+' ? 0 : 1);    // bam2gtf.c:35-37
                a.rows.start[row] = S.start[tid]; a.rows.end[row] = S.end[tid];
                a.rows.ex_beg[row] = ex_base + ex_excl; a.rows.ex_n[row] = my_ex;
            }
        }
    }
    if (!do_exon || ex_total == 0) return;
    if ((int64_t)ex_base + ex_total > a.ex.cap) return;          // host re-runs with a larger pool
    if (!ovf) {
        if (tid < nr && my_ex) {
            int *es = a.ex.es + ex_base + ex_excl, *ee = a.ex.ee + ex_base + ex_excl;
            if (S.off[tid + 1] == S.off[tid]) { es[0] = S.pos[tid] + 1; ee[0] = S.pos[tid]; }   // no ops at all: the open exon (bam2gtf.c:74-76)
            else {
                const int *xs = S.es[S.wof[tid]] + S.sbeg[tid], *xe = S.ee[S.wof[tid]] + S.sbeg[tid];
                for (uint32_t k = 0; k < my_ex; ++k) { es[k] = xs[k]; ee[k] = xe[k]; }
            }
        }
    } else {
        S.cnt[tid] = (int)ex_excl;
        __syncthreads();
        for (int li = w; li < nr; li += ST_WARPS) {
            const int64_t rr = r0 + li;
            if (!(S.mask[li] && !(a.b.flag[rr] & 4))) continue;
            const uint32_t *c = a.b.cigar + a.b.cigar_off[rr]; const int n_c = (int)(a.b.cigar_off[rr + 1] - a.b.cigar_off[rr]);
            WalkStats ws; walk_warp<true>(c, n_c, a.b.pos[rr], a.ep, a.ex.es + ex_base + S.cnt[li], a.ex.ee + ex_base + S.cnt[li], ws);
        }
    }
}

int stream_reads_per_tile() { return ST_R; }

void launch_cigar_scan(const ScanArgs &a, int n_tiles, bool stream_mode, size_t smem_bytes, cudaStream_t st)
{
    if (n_tiles <= 0) return;
    if (stream_mode) {
        static bool attr = false;
        if (!attr) { cudaFuncSetAttribute(cigar_stream_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(StreamSmem)); attr = true; }
        cigar_stream_kernel<<<n_tiles, ST_THREADS, sizeof(StreamSmem), st>>>(a);
    } else {
        cudaFuncSetAttribute(cigar_scan_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes);
        cigar_scan_kernel<<<n_tiles, SCAN_THREADS, smem_bytes, st>>>(a);
    }
    LRB_COUNT_LAUNCH();
}

// ---------------------------------------------------------------------------------------- qname-run selection
// Rows are the passing records in input order, so a run of equal qname hashes over adjacent rows is exactly the run the
// reference's state machine sees (bam_filter.c:133-153).  The head row of each run replays the machine.
__global__ void select_runs_kernel(const uint64_t *__restrict__ qhash, const uint32_t *__restrict__ row_read, int64_t n_rows,
                                   const int32_t *__restrict__ score, const int32_t *__restrict__ intron_n,
                                   lrb_filter_params fp, uint8_t *keep_row_mask, uint8_t *keep_rec_mask)
{
    int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n_rows) return;
    uint32_t r = row_read[k];
    uint64_t h = qhash[r];
    if (k > 0 && qhash[row_read[k - 1]] == h) return;          // not a run head
    int b_score = score[r], s_score = 0, b_intron = intron_n[r]; int64_t best = k;
    for (int64_t j = k + 1; j < n_rows; ++j) {
        uint32_t rj = row_read[j];
        if (qhash[rj] != h) break;
        int sc = score[rj];
        if (sc > b_score) { best = j; s_score = b_score; b_score = sc; b_intron = intron_n[rj]; }
        else if (sc > s_score) s_score = sc;
    }
    if ((float)s_score < __fmul_rn(fp.sec_rat, (float)b_score) && b_intron >= fp.min_intron_n) {
        keep_row_mask[best] = 1;
        keep_rec_mask[row_read[best]] = 1;
    }
}

void launch_select_runs(const DBatch &b, const uint32_t *row_read, int64_t n_rows, const int32_t *score, const int32_t *intron_n,
                        lrb_filter_params fp, uint8_t *keep_row_mask, uint8_t *keep_rec_mask, cudaStream_t st)
{
    if (n_rows <= 0) return;
    int th = 256; int64_t bl = (n_rows + th - 1) / th;
    select_runs_kernel<<<(unsigned)bl, th, 0, st>>>(b.qhash, row_read, n_rows, score, intron_n, fp, keep_row_mask, keep_rec_mask);
    LRB_COUNT_LAUNCH();
}

// The same state machine on the record stream itself (fused filter + exon pass): the passing subsequence is walked through
// the pass mask, so no compacted row list -- and no host round trip for its size -- is needed in front of the selection.
__global__ void select_records_kernel(const uint64_t *__restrict__ qhash, const uint8_t *__restrict__ pass, int64_t n,
                                      const int32_t *__restrict__ score, const int32_t *__restrict__ intron_n, lrb_filter_params fp, uint8_t *keep_rec_mask)
{
    int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n || !pass[r]) return;
    const uint64_t h = qhash[r];
    int64_t p = r - 1;
    while (p >= 0 && !pass[p]) --p;
    if (p >= 0 && qhash[p] == h) return;                        // not a run head
    int b_score = score[r], s_score = 0, b_intron = intron_n[r]; int64_t best = r;
    for (int64_t j = r + 1; j < n; ++j) {
        if (!pass[j]) continue;
        if (qhash[j] != h) break;
        int sc = score[j];
        if (sc > b_score) { best = j; s_score = b_score; b_score = sc; b_intron = intron_n[j]; }
        else if (sc > s_score) s_score = sc;
    }
    if ((float)s_score < __fmul_rn(fp.sec_rat, (float)b_score) && b_intron >= fp.min_intron_n) keep_rec_mask[best] = 1;
}
void launch_select_records(const DBatch &b, const uint8_t *pass, const int32_t *score, const int32_t *intron_n, lrb_filter_params fp,
                           uint8_t *keep_rec_mask, cudaStream_t st)
{
    if (b.n <= 0) return;
    select_records_kernel<<<(unsigned)((b.n + 255) / 256), 256, 0, st>>>(b.qhash, pass, b.n, score, intron_n, fp, keep_rec_mask);
    LRB_COUNT_LAUNCH();
}

// ordered compaction of the kept records + gather of their (record-indexed) rows into the compact row table
__global__ void __launch_bounds__(256) compact_gather_kernel(const uint8_t *__restrict__ mask, int64_t n, DRows src, DRows dst, uint32_t *keep_idx,
                                                             uint64_t *tile_state, uint32_t *ticket, uint64_t *total)
{
    constexpr int ITEMS = 8;
    __shared__ uint32_t s_scan[33];
    __shared__ uint32_t s_tile; __shared__ uint64_t s_excl;
    if (threadIdx.x == 0) s_tile = atomicAdd(ticket, 1u);
    __syncthreads();
    const int tile = (int)s_tile;
    const int64_t base = ((int64_t)tile * 256 + threadIdx.x) * ITEMS;
    uint32_t cnt = 0, m = 0;
    if (base + ITEMS <= n) {                                     // 8 mask bytes in one load (base is a multiple of 8)
        const uint2 q = *(const uint2 *)(mask + base);
#pragma unroll
        for (int i = 0; i < 4; ++i) { if ((q.x >> (8 * i)) & 0xffu) m |= 1u << i; if ((q.y >> (8 * i)) & 0xffu) m |= 1u << (4 + i); }
    } else {
        for (int i = 0; i < ITEMS; ++i) if (base + i < n && mask[base + i]) m |= 1u << i;
    }
    cnt = __popc(m);
    uint32_t tot, excl = block_excl_sum(cnt, s_scan, &tot);
    if (warp_id() == 0) { uint64_t e = lookback_exclusive(tile_state, tile, tot, OpAdd()); if (lane_id() == 0) s_excl = e; }
    __syncthreads();
    uint32_t o = (uint32_t)s_excl + excl;
    while (m) {
        const int i = __ffs(m) - 1; m &= m - 1;
        const uint32_t s = (uint32_t)(base + i);
        keep_idx[o] = s;
        dst.read_idx[o] = s; dst.tid[o] = src.tid[s]; dst.start[o] = src.start[s]; dst.end[o] = src.end[s];
        dst.is_rev[o] = src.is_rev[s]; dst.ex_beg[o] = src.ex_beg[s]; dst.ex_n[o] = src.ex_n[s];
        ++o;
    }
    if ((int64_t)(tile + 1) * 256 * ITEMS >= n && threadIdx.x == 0) *total = s_excl + tot;
}
void launch_compact_gather(const uint8_t *mask, int64_t n, const DRows &src, DRows &dst, uint32_t *keep_idx,
                           uint64_t *tile_state, uint32_t *ticket, uint64_t *total, cudaStream_t st)
{
    if (n <= 0) { cudaMemsetAsync(total, 0, 8, st); return; }
    int64_t per = 256 * 8, bl = (n + per - 1) / per;
    cudaMemsetAsync(tile_state, 0, (size_t)bl * 8, st); cudaMemsetAsync(ticket, 0, 4, st);
    compact_gather_kernel<<<(unsigned)bl, 256, 0, st>>>(mask, n, src, dst, keep_idx, tile_state, ticket, total);
    LRB_COUNT_LAUNCH();
}

// -------------------------------------------------------------------------------------------- mask compaction
static constexpr int CM_THREADS = 256, CM_ITEMS = 8;
__global__ void __launch_bounds__(CM_THREADS) compact_mask_kernel(const uint8_t *__restrict__ mask, int64_t n, const uint32_t *__restrict__ map,
                                                                  uint32_t *out, uint32_t *out2, uint64_t *tile_state, uint32_t *ticket, uint64_t *total)
{
    __shared__ uint32_t s_scan[33];
    __shared__ uint32_t s_tile; __shared__ uint64_t s_excl;
    if (threadIdx.x == 0) s_tile = atomicAdd(ticket, 1u);
    __syncthreads();
    const int tile = (int)s_tile;
    const int64_t base = ((int64_t)tile * CM_THREADS + threadIdx.x) * CM_ITEMS;
    uint32_t cnt = 0; uint8_t m[CM_ITEMS];
#pragma unroll
    for (int i = 0; i < CM_ITEMS; ++i) { m[i] = (base + i < n) ? mask[base + i] : 0; cnt += m[i] != 0; }
    uint32_t tot, excl = block_excl_sum(cnt, s_scan, &tot);
    if (warp_id() == 0) { uint64_t e = lookback_exclusive(tile_state, tile, tot, OpAdd()); if (lane_id() == 0) s_excl = e; }
    __syncthreads();
    uint32_t o = (uint32_t)s_excl + excl;
#pragma unroll
    for (int i = 0; i < CM_ITEMS; ++i)
        if (m[i]) { uint32_t idx = (uint32_t)(base + i); out[o] = map ? map[idx] : idx; if (out2) out2[o] = idx; ++o; }
    if ((int64_t)(tile + 1) * CM_THREADS * CM_ITEMS >= n && threadIdx.x == 0) *total = s_excl + tot;
}

void launch_compact_mask(const uint8_t *mask, int64_t n, const uint32_t *map, uint32_t *out, uint32_t *out2,
                         uint64_t *tile_state, uint32_t *ticket, uint64_t *total, cudaStream_t st)
{
    if (n <= 0) { cudaMemsetAsync(total, 0, 8, st); return; }
    int64_t per = (int64_t)CM_THREADS * CM_ITEMS, bl = (n + per - 1) / per;
    cudaMemsetAsync(tile_state, 0, (size_t)bl * 8, st); cudaMemsetAsync(ticket, 0, 4, st);
    compact_mask_kernel<<<(unsigned)bl, CM_THREADS, 0, st>>>(mask, n, map, out, out2, tile_state, ticket, total);
    LRB_COUNT_LAUNCH();
}

// ------------------------------------------------------------------------------------------------ row gather
__global__ void gather_rows_kernel(DRows src, const uint32_t *__restrict__ sel, int64_t n_sel, DRows dst)
{
    int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n_sel) return;
    uint32_t s = sel[k];
    dst.read_idx[k] = src.read_idx[s]; dst.tid[k] = src.tid[s]; dst.start[k] = src.start[s]; dst.end[k] = src.end[s];
    dst.is_rev[k] = src.is_rev[s]; dst.ex_beg[k] = src.ex_beg[s]; dst.ex_n[k] = src.ex_n[s];
}
void launch_gather_rows(const DRows &src, const uint32_t *sel, int64_t n_sel, DRows &dst, cudaStream_t st)
{
    if (n_sel <= 0) return;
    int th = 256; int64_t bl = (n_sel + th - 1) / th;
    gather_rows_kernel<<<(unsigned)bl, th, 0, st>>>(src, sel, n_sel, dst);
    LRB_COUNT_LAUNCH();
}

// ----------------------------------------------------------------------------------------------- generic scans
static constexpr int GS_THREADS = 256, GS_ITEMS = 8;
__global__ void __launch_bounds__(GS_THREADS) scan_max_u64_kernel(uint64_t *data, int64_t n, uint64_t *tile_state, uint32_t *ticket)
{
    __shared__ uint64_t s_w[GS_THREADS / 32]; __shared__ uint32_t s_tile; __shared__ uint64_t s_excl;
    if (threadIdx.x == 0) s_tile = atomicAdd(ticket, 1u);
    __syncthreads();
    const int tile = (int)s_tile, lane = lane_id(), w = warp_id();
    const int64_t base = ((int64_t)tile * GS_THREADS + threadIdx.x) * GS_ITEMS;
    uint64_t v[GS_ITEMS], run = 0;
#pragma unroll
    for (int i = 0; i < GS_ITEMS; ++i) { v[i] = (base + i < n) ? data[base + i] : 0; run = v[i] > run ? v[i] : run; v[i] = run; }
    uint64_t inc = run;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { uint64_t t = __shfl_up_sync(FULL, inc, o); if (lane >= o && t > inc) inc = t; }
    if (lane == 31) s_w[w] = inc;
    __syncthreads();
    uint64_t pre = 0;                                            // max over warps before mine
    for (int k = 0; k < w; ++k) pre = s_w[k] > pre ? s_w[k] : pre;
    uint64_t tot = 0;
    for (int k = 0; k < GS_THREADS / 32; ++k) tot = s_w[k] > tot ? s_w[k] : tot;
    uint64_t left = __shfl_up_sync(FULL, inc, 1); if (lane == 0) left = 0;
    uint64_t tpre = left > pre ? left : pre;                     // exclusive prefix max of this thread inside the tile
    if (w == 0) { uint64_t e = lookback_exclusive(tile_state, tile, tot, OpMax()); if (lane == 0) s_excl = e; }
    __syncthreads();
    uint64_t ex = s_excl > tpre ? s_excl : tpre;
#pragma unroll
    for (int i = 0; i < GS_ITEMS; ++i) if (base + i < n) data[base + i] = v[i] > ex ? v[i] : ex;
}
void launch_scan_max_u64(uint64_t *data, int64_t n, uint64_t *tile_state, uint32_t *ticket, cudaStream_t st)
{
    if (n <= 0) return;
    int64_t per = (int64_t)GS_THREADS * GS_ITEMS, bl = (n + per - 1) / per;
    cudaMemsetAsync(tile_state, 0, (size_t)bl * 8, st); cudaMemsetAsync(ticket, 0, 4, st);
    scan_max_u64_kernel<<<(unsigned)bl, GS_THREADS, 0, st>>>(data, n, tile_state, ticket);
    LRB_COUNT_LAUNCH();
}

__global__ void __launch_bounds__(GS_THREADS) scan_sum_u32_kernel(const uint32_t *__restrict__ in, uint32_t *out, int64_t n,
                                                                  uint64_t *tile_state, uint32_t *ticket, uint64_t *total)
{
    __shared__ uint32_t s_scan[33]; __shared__ uint32_t s_tile; __shared__ uint64_t s_excl;
    if (threadIdx.x == 0) s_tile = atomicAdd(ticket, 1u);
    __syncthreads();
    const int tile = (int)s_tile;
    const int64_t base = ((int64_t)tile * GS_THREADS + threadIdx.x) * GS_ITEMS;
    uint32_t v[GS_ITEMS], sum = 0;
#pragma unroll
    for (int i = 0; i < GS_ITEMS; ++i) { v[i] = (base + i < n) ? in[base + i] : 0; sum += v[i]; }
    uint32_t tot, excl = block_excl_sum(sum, s_scan, &tot);
    if (warp_id() == 0) { uint64_t e = lookback_exclusive(tile_state, tile, tot, OpAdd()); if (lane_id() == 0) s_excl = e; }
    __syncthreads();
    uint32_t o = (uint32_t)s_excl + excl;
#pragma unroll
    for (int i = 0; i < GS_ITEMS; ++i) if (base + i < n) { out[base + i] = o; o += v[i]; }
    if ((int64_t)(tile + 1) * GS_THREADS * GS_ITEMS >= n && threadIdx.x == 0 && total) *total = s_excl + tot;
}
void launch_scan_sum_u32(const uint32_t *in, uint32_t *out_excl, int64_t n, uint64_t *tile_state, uint32_t *ticket, uint64_t *total, cudaStream_t st)
{
    if (n <= 0) { if (total) cudaMemsetAsync(total, 0, 8, st); return; }
    int64_t per = (int64_t)GS_THREADS * GS_ITEMS, bl = (n + per - 1) / per;
    cudaMemsetAsync(tile_state, 0, (size_t)bl * 8, st); cudaMemsetAsync(ticket, 0, 4, st);
    scan_sum_u32_kernel<<<(unsigned)bl, GS_THREADS, 0, st>>>(in, out_excl, n, tile_state, ticket, total);
    LRB_COUNT_LAUNCH();
}

}  // namespace lrbk
