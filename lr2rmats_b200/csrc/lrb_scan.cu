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
static constexpr int EX_STAGE = 3072;             // exon staging slots per tile (24 KB for starts+ends)

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

// ---- flat mode (long CIGARs): the staged words of a tile are ONE array of ops, 256 threads take equal slices of it whatever
// the read boundaries are, and everything a walk carries from op to op becomes a block-wide scan:
//   reference bases consumed so far   exclusive sum over the slices; a read's own origin = the value at its first op
//   the latest cut op and the reference position behind it   exclusive max-scan of (op index, position) pairs
//   exons emitted so far              exclusive sum of the emit flags; a read's exon k = emitted cuts since its first op
// so a read of 2000 ops costs each of 256 threads a few ops instead of one warp 63 rounds of shuffles.
template <int WM> struct FlatSmem { };
template <> struct FlatSmem<2> {
    int off[SCAN_THREADS + 1], pos[SCAN_THREADS];
    uint32_t base[SCAN_THREADS], pend[SCAN_THREADS], intron[SCAN_THREADS], del[SCAN_THREADS], EB[SCAN_THREADS], EE[SCAN_THREADS];
    uint16_t ebl[SCAN_THREADS], ebo[SCAN_THREADS], eel[SCAN_THREADS], eeo[SCAN_THREADS];
    uint32_t tE[SCAN_THREADS]; unsigned long long w64[SCAN_THREADS / 32 + 1];
    uint8_t emit[SCAN_THREADS];
};

// WM: 0 one thread per read (short CIGARs), 1 one warp per read, 2 flat (tiles that do not fit the stage fall back to 1)
// WM 3 = flat mode compiled for 6 resident CTAs per SM (40 registers) instead of 5
template <int WMX>
__global__ void __launch_bounds__(SCAN_THREADS, WMX == 3 ? 6 : 0) cigar_scan_kernel(ScanArgs a)
{
    constexpr int WM = WMX == 3 ? 2 : WMX;
    constexpr bool WARP_MODE = WM != 0;
    __shared__ FlatSmem<WM> fs;
    extern __shared__ __align__(16) uint32_t smem[];
    uint32_t *s_words = smem;                                   // stage_words
    int *s_es = (int *)(smem + a.stage_words), *s_ee = s_es + EX_STAGE;
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
    const uint32_t w_lo = a.b.cigar_off[r0], w_hi = a.b.cigar_off[r1];
    const uint32_t nw = w_hi - w_lo;
    const bool staged = nw <= (uint32_t)a.stage_words;
    const bool do_filter = a.mode != 1, do_exon = a.mode != 0;
    if (staged) {
        // coalesced stage: scalar head up to 16-byte alignment, then 128-bit streaming loads
        const uint32_t *src = a.b.cigar + w_lo;
        uint32_t head = (uint32_t)((4 - (w_lo & 3)) & 3); if (head > nw) head = nw;
        if ((uint32_t)tid < head) s_words[tid] = ldg_stream_u32(src + tid);
        uint32_t nv = (nw - head) >> 2;
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
        uint32_t off = a.b.cigar_off[r];
        return staged ? (s_words + (off - w_lo)) : (a.b.cigar + off);
    };
    auto finish_read = [&](int li, int64_t r, uint32_t c0, uint32_t c1, int n_c, const WalkStats &w) {
        bool mask;
        if (do_filter) {
            int sc = 0; bool p = filter_pass(a, r, c0, c1, n_c, w, &sc);
            a.pass[r] = p ? 1 : 0;
            if (p) { a.score[r] = sc; a.intron_n[r] = w.intron_n; }
            mask = p;
        } else mask = a.sel_mask ? (a.sel_mask[r] != 0) : true;
        bool unmapped = (a.b.flag[r] & 4) != 0;
        s_mask[li] = mask ? 1 : 0;
        s_cnt[li] = (mask && do_exon && !unmapped) ? w.n_exon : 0;
        s_start[li] = w.first_start; s_end[li] = w.last_end;
    };
    bool parked = false, park_ovf = false; int park_last_start = 0;      // thread mode: exons parked in the staged words by walk 1
    // flat mode state of this thread's slice [fa0, fa1) of the tile's ops
    bool flat = false; int fa0 = 0, fa1 = 0, fli = 0; uint32_t fP = 0, fE = 0; int fcut = -1; uint32_t fcutP = 0;
    // the ops are read from the staged copy when the tile fits the stage, else straight from the pool (the slices are contiguous:
    // every pass streams them again through L1 / L2), so a flat tile is not bounded by shared memory
    const uint32_t *f_words = staged ? s_words : a.b.cigar + w_lo;
    if constexpr (WM == 2) flat = nr <= SCAN_THREADS;
    if constexpr (WM == 2) if (flat) {
        if (tid < nr) fs.off[tid] = (int)(a.b.cigar_off[r0 + tid] - w_lo);
        if (tid == 0) fs.off[nr] = (int)nw;
        if (tid < nr) { fs.pos[tid] = a.b.pos[r0 + tid]; fs.intron[tid] = 0; fs.del[tid] = 0; fs.base[tid] = 0; fs.pend[tid] = 0; fs.ebl[tid] = fs.ebo[tid] = fs.eel[tid] = fs.eeo[tid] = 0; }
        __syncthreads();
        const int K = (int)((nw + SCAN_THREADS - 1) / SCAN_THREADS);
        fa0 = min((int)nw, tid * K); fa1 = min((int)nw, fa0 + K);
        {   // read of the first op of the slice: first r with off[r + 1] > fa0
            int lo = 0, hi = nr;
            while (lo < hi) { const int m = (lo + hi) >> 1; if (fs.off[m + 1] > fa0) hi = m; else lo = m + 1; }
            fli = lo;
        }
        // pass 1: reference bases of the slice -> exclusive sum
        uint32_t tot = 0;
        for (int i = fa0; i < fa1; ++i) { const uint32_t x = f_words[i]; if (op_ref(x & 15u)) tot += x >> 4; }
        uint32_t all; fP = block_excl_sum(tot, s_scan, &all);
        // pass 2: read origins, per-read filter statistics, the slice's last cut
        {
            uint32_t P = fP; int li = fli; uint32_t intr = 0, dl = 0; unsigned long long lastcut = 0;
            for (int i = fa0; i < fa1; ++i) {
                while (i >= fs.off[li + 1]) { if (intr) atomicAdd(&fs.intron[li], intr); if (dl) atomicAdd(&fs.del[li], dl); intr = dl = 0; ++li; }
                if (i == fs.off[li]) fs.base[li] = P;
                const uint32_t x = f_words[i], l = x >> 4, op = x & 15u;
                if (op == OP_N) ++intr; else if (op == OP_D) dl += l;
                const bool cut = (op == OP_N && (int)l >= a.ep.min_intron) || (op == OP_D && (int)l > a.ep.max_delet);
                if (op_ref(op)) P += l;
                if (cut) lastcut = ((unsigned long long)(uint32_t)(i + 1) << 32) | P;
                if (i == fs.off[li + 1] - 1) fs.pend[li] = P;
            }
            if (fa0 < fa1) { if (intr) atomicAdd(&fs.intron[li], intr); if (dl) atomicAdd(&fs.del[li], dl); }
            // exclusive max-scan of (cut index + 1, position behind the cut) over the slices
            unsigned long long inc = lastcut;
            const int lane = lane_id(), w = warp_id();
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const unsigned long long y = __shfl_up_sync(FULL, inc, o); if (lane >= o && y > inc) inc = y; }
            if (lane == 31) fs.w64[w] = inc;
            __syncthreads();
            unsigned long long pre = 0;
            for (int k = 0; k < w; ++k) pre = fs.w64[k] > pre ? fs.w64[k] : pre;
            unsigned long long left = __shfl_up_sync(FULL, inc, 1); if (lane == 0) left = 0;
            const unsigned long long ex = left > pre ? left : pre;
            fcut = (int)(ex >> 32) - 1; fcutP = (uint32_t)ex;
        }
        __syncthreads();
        // pass 3: emit flags -> per-slice counts, per-read (owner slice, local count) at the first and behind the last op
        {
            uint32_t P = fP; int li = fli, lc = fcut; uint32_t lcP = fcutP, ec = 0;
            for (int i = fa0; i < fa1; ++i) {
                while (i >= fs.off[li + 1]) ++li;
                if (i == fs.off[li]) { fs.ebl[li] = (uint16_t)ec; fs.ebo[li] = (uint16_t)tid; }
                const uint32_t x = f_words[i], l = x >> 4, op = x & 15u;
                const bool cut = (op == OP_N && (int)l >= a.ep.min_intron) || (op == OP_D && (int)l > a.ep.max_delet);
                if (cut) {
                    const bool first_cut = lc < fs.off[li];
                    const int end_before = fs.pos[li] + (int)(P - fs.base[li]);
                    const int start = first_cut ? fs.pos[li] + 1 : fs.pos[li] + (int)(lcP - fs.base[li]) + 1;
                    if (first_cut || end_before - start + 1 >= a.ep.min_exon) ++ec;
                }
                if (op_ref(op)) P += l;
                if (cut) { lc = i; lcP = P; }
                if (i == fs.off[li + 1] - 1) { fs.eel[li] = (uint16_t)ec; fs.eeo[li] = (uint16_t)tid; }
            }
            uint32_t eall; fE = block_excl_sum(ec, s_scan, &eall);
            fs.tE[tid] = fE;
        }
        __syncthreads();
        if (tid < nr) {
            const int64_t r = r0 + tid;
            const int n_c = fs.off[tid + 1] - fs.off[tid];
            const uint32_t c0 = n_c > 0 ? f_words[fs.off[tid]] : 0u, c1 = n_c > 0 ? f_words[fs.off[tid + 1] - 1] : 0u;
            WalkStats w;
            fs.EB[tid] = n_c > 0 ? fs.tE[fs.ebo[tid]] + fs.ebl[tid] : 0u;
            fs.EE[tid] = n_c > 0 ? fs.tE[fs.eeo[tid]] + fs.eel[tid] : 0u;
            w.n_exon = (int)(fs.EE[tid] - fs.EB[tid]) + 1; w.intron_n = (int)fs.intron[tid]; w.del_len = (int)fs.del[tid];
            w.ref_len = n_c > 0 ? (int)(fs.pend[tid] - fs.base[tid]) : 0; w.first_start = fs.pos[tid] + 1; w.last_end = fs.pos[tid] + w.ref_len;
            finish_read(tid, r, c0, c1, n_c, w);
        }
    }
    if (flat) { }
    else if (!WARP_MODE) {
        if (tid < nr) {
            int64_t r = r0 + tid;
            const uint32_t *c = read_ptr(r); int n_c = (int)(a.b.cigar_off[r + 1] - a.b.cigar_off[r]);
            const uint32_t c0 = n_c > 0 ? c[0] : 0u, c1 = n_c > 0 ? c[n_c - 1] : 0u;
            WalkStats w;
            if (staged && do_exon) { parked = true; walk_seq_inplace(const_cast<uint32_t *>(c), n_c, a.b.pos[r], a.ep, w, &park_ovf, &park_last_start); }
            else walk_seq<false>(c, n_c, a.b.pos[r], a.ep, nullptr, nullptr, w);
            finish_read(tid, r, c0, c1, n_c, w);
        }
    } else {
        for (int li = warp_id(); li < nr; li += SCAN_THREADS / 32) {
            int64_t r = r0 + li;
            const uint32_t *c = read_ptr(r); int n_c = (int)(a.b.cigar_off[r + 1] - a.b.cigar_off[r]);
            WalkStats w; walk_warp<false>(c, n_c, a.b.pos[r], a.ep, nullptr, nullptr, w);
            if (lane_id() == 0) finish_read(li, r, n_c > 0 ? c[0] : 0u, n_c > 0 ? c[n_c - 1] : 0u, n_c, w);
        }
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

    // ---- walk 2: emit exons (into the shared staging buffer when the tile fits, else straight to HBM)
    // warp mode stages the exons of the tile and writes them out coalesced; thread mode moves each read's parked exons
    // straight to the pools (one staging buffer less: 6 instead of 4 resident CTAs per SM cover the look-back waits)
    const bool ex_staged = WARP_MODE && a.stage_words > 0 && ex_total <= (uint32_t)EX_STAGE;
    s_cnt[tid] = (int)ex_excl;                                   // reuse as local exon offset (own slot only)
    if constexpr (WM == 2) if (flat && tid < nr) fs.emit[tid] = my_ex > 0;
    __syncthreads();
    const bool room = (int64_t)ex_base + ex_total <= a.ex.cap;   // host re-runs with a larger pool otherwise
    if (!room) return;
    if constexpr (WM == 2) if (flat) {
        // every slice replays its ops once more, now with the exon slot of each emitted cut: slot of the read + cuts emitted since its first op
        int *bes = ex_staged ? s_es : a.ex.es + ex_base, *bee = ex_staged ? s_ee : a.ex.ee + ex_base;
        uint32_t P = fP, E = fE; int li = fli, lc = fcut; uint32_t lcP = fcutP;
        for (int i = fa0; i < fa1; ++i) {
            while (i >= fs.off[li + 1]) ++li;
            const uint32_t x = f_words[i], l = x >> 4, op = x & 15u;
            const bool cut = (op == OP_N && (int)l >= a.ep.min_intron) || (op == OP_D && (int)l > a.ep.max_delet);
            if (cut) {
                const bool first_cut = lc < fs.off[li];
                const int end_before = fs.pos[li] + (int)(P - fs.base[li]);
                const int start = first_cut ? fs.pos[li] + 1 : fs.pos[li] + (int)(lcP - fs.base[li]) + 1;
                if (first_cut || end_before - start + 1 >= a.ep.min_exon) {
                    if (fs.emit[li]) { const uint32_t k = (uint32_t)s_cnt[li] + (E - fs.EB[li]); bes[k] = start; bee[k] = end_before; }
                    ++E;
                }
            }
            if (op_ref(op)) P += l;
            if (cut) { lc = i; lcP = P; }
            if (i == fs.off[li + 1] - 1 && fs.emit[li]) {           // the open exon behind the last op (bam2gtf.c:74-76)
                const uint32_t k = (uint32_t)s_cnt[li] + (fs.EE[li] - fs.EB[li]);
                bes[k] = lc >= fs.off[li] ? fs.pos[li] + (int)(lcP - fs.base[li]) + 1 : fs.pos[li] + 1;
                bee[k] = fs.pos[li] + (int)(P - fs.base[li]);
            }
        }
        if (tid < nr && fs.emit[tid] && fs.off[tid + 1] == fs.off[tid]) { bes[s_cnt[tid]] = fs.pos[tid] + 1; bee[s_cnt[tid]] = fs.pos[tid]; }   // no ops at all
    }
    if (flat) { }
    else if (!WARP_MODE) {
        if (tid < nr && my_ex) {
            int64_t r = r0 + tid;
            const uint32_t *c = read_ptr(r); int n_c = (int)(a.b.cigar_off[r + 1] - a.b.cigar_off[r]);
            int *es = ex_staged ? s_es + ex_excl : a.ex.es + ex_base + ex_excl;
            int *ee = ex_staged ? s_ee + ex_excl : a.ex.ee + ex_base + ex_excl;
            if (parked && !park_ovf) {                            // move the parked exons; the open one is still in registers
                const int last = (int)my_ex - 1;
                for (int k = 0; k < last; ++k) { es[k] = (int)c[2 * k]; ee[k] = (int)c[2 * k + 1]; }
                es[last] = park_last_start; ee[last] = s_end[tid];
            } else {
                if (parked) c = a.b.cigar + a.b.cigar_off[r];     // the staged copy is partly overwritten: walk the pool
                WalkStats w; walk_seq<true>(c, n_c, a.b.pos[r], a.ep, es, ee, w);
            }
        }
    } else {
        for (int li = warp_id(); li < nr; li += SCAN_THREADS / 32) {
            // my_ex of read li lives in thread li's registers: fetch count via the row test below
            int64_t r = r0 + li;
            bool walk = s_mask[li] && !(a.b.flag[r] & 4);
            if (!walk) continue;
            const uint32_t *c = read_ptr(r); int n_c = (int)(a.b.cigar_off[r + 1] - a.b.cigar_off[r]);
            uint32_t lo = (uint32_t)s_cnt[li];
            int *es = ex_staged ? s_es + lo : a.ex.es + ex_base + lo;
            int *ee = ex_staged ? s_ee + lo : a.ex.ee + ex_base + lo;
            WalkStats w; walk_warp<true>(c, n_c, a.b.pos[r], a.ep, es, ee, w);
        }
    }
    if (ex_staged) {
        __syncthreads();
        for (uint32_t i = tid; i < ex_total; i += SCAN_THREADS) { a.ex.es[ex_base + i] = s_es[i]; a.ex.ee[ex_base + i] = s_ee[i]; }
    }
}

void launch_cigar_scan(const ScanArgs &a, int n_tiles, bool warp_mode, size_t smem_bytes, cudaStream_t st)
{
    if (n_tiles <= 0) return;
    static int flat = -1;
    if (flat < 0) { const char *e = getenv("LRB_SCAN_FLAT"); flat = e ? atoi(e) : 1; }
    static int occ6 = -1;
    if (occ6 < 0) { const char *e = getenv("LRB_SCAN_OCC6"); occ6 = e ? atoi(e) : 1; }
    if (warp_mode && flat && occ6) {
        cudaFuncSetAttribute(cigar_scan_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes);
        cigar_scan_kernel<3><<<n_tiles, SCAN_THREADS, smem_bytes, st>>>(a);
    } else if (warp_mode && flat) {
        cudaFuncSetAttribute(cigar_scan_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes);
        cigar_scan_kernel<2><<<n_tiles, SCAN_THREADS, smem_bytes, st>>>(a);
    } else if (warp_mode) {
        cudaFuncSetAttribute(cigar_scan_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes);
        cigar_scan_kernel<1><<<n_tiles, SCAN_THREADS, smem_bytes, st>>>(a);
    } else {
        cudaFuncSetAttribute(cigar_scan_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes);
        cigar_scan_kernel<0><<<n_tiles, SCAN_THREADS, smem_bytes, st>>>(a);
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
