// lrb_sort.cu -- stable LSD radix sort of the kept rows by coordinate, on the device.
//
// The pipeline sorts between `filter` and `update-gtf` with an external tool (Snakefile:90: `lr2rmats filter ... |
// samtools sort > filtered.bam`), because update_gtf's sweep needs (tid,start)-sorted input (update_gtf.c:41).  With the
// rows already in HBM that hop is a key-value sort of ~25 bytes per row: the key is samtools' coordinate key
//     tid << 32 | (pos + 1) << 1 | reverse-strand flag             (bam_sort.c, bam1_lt; stable for equal keys)
// and the value the row index; the rows are then permuted once (their exon chains stay where they are: rows carry
// ex_beg / ex_n).  Hand-written, no CUB: 8-bit digits, per pass  histogram -> exclusive scan -> stable scatter; only the
// passes the largest key needs are run.
//   sort_keys_kernel      key + identity permutation per row, block-reduced maximum key
//   sort_hist_kernel      digit counts per 2048-row tile, digit-major (hist[d * n_tiles + tile])
//   sort_scan_kernel      exclusive scan of that matrix (one CTA: <= 256 * n_tiles counters)
//   sort_scatter_kernel   a warp owns 256 consecutive rows and ranks them in rounds of 32 with __match_any_sync (rank =
//                         earlier rows of the tile with the same digit), so equal digits keep their order
//   rows_permute_kernel   gathers the seven row fields through the sorted permutation
#include "lrb_common.cuh"
#include "lrb_kernels.cuh"

namespace lrbk {

extern int64_t g_launches_sort;
int64_t g_launches_sort = 0;
#define LRB_COUNT_LAUNCH() (++g_launches_sort)

static constexpr int RS_THREADS = 256, RS_TILE = 2048, RS_WARPS = RS_THREADS / 32, RS_ROUNDS = RS_TILE / RS_THREADS;   // a warp owns RS_ROUNDS * 32 consecutive rows

__global__ void __launch_bounds__(RS_THREADS) sort_keys_kernel(DRows rows, const uint16_t *__restrict__ flag, uint64_t *__restrict__ keys,
                                                               uint32_t *__restrict__ idx, unsigned long long *max_key)
{
    const int64_t i = (int64_t)blockIdx.x * RS_THREADS + threadIdx.x;
    uint64_t k = 0;
    if (i < rows.n) {
        // rows.start = pos + 1 (bam2gtf.c:45); the strand bit is the record's FLAG 0x10, not the XS-derived transcript strand
        const uint32_t rev = flag ? ((flag[rows.read_idx[i]] & 16) ? 1u : 0u) : (uint32_t)rows.is_rev[i];
        k = ((uint64_t)(uint32_t)rows.tid[i] << 32) | ((uint64_t)(uint32_t)rows.start[i] << 1) | rev;
        keys[i] = k; idx[i] = (uint32_t)i;
    }
    uint64_t m = k;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { const uint64_t t = __shfl_xor_sync(FULL, m, o); m = t > m ? t : m; }
    if (lane_id() == 0 && m) atomicMax(max_key, (unsigned long long)m);
}

__global__ void __launch_bounds__(RS_THREADS) sort_hist_kernel(const uint64_t *__restrict__ keys, int64_t n, int shift, uint32_t *__restrict__ hist, int n_tiles)
{
    __shared__ uint32_t s_h[256];
    s_h[threadIdx.x] = 0;
    __syncthreads();
    const int64_t base = (int64_t)blockIdx.x * RS_TILE;
#pragma unroll
    for (int r = 0; r < RS_TILE / RS_THREADS; ++r) {
        const int64_t i = base + r * RS_THREADS + threadIdx.x;
        if (i < n) atomicAdd(&s_h[(keys[i] >> shift) & 255u], 1u);
    }
    __syncthreads();
    hist[(size_t)threadIdx.x * n_tiles + blockIdx.x] = s_h[threadIdx.x];
}

// exclusive scan of m counters in place, one CTA of 1024 threads: every thread owns a contiguous slice
__global__ void __launch_bounds__(1024) sort_scan_kernel(uint32_t *__restrict__ hist, int64_t m)
{
    __shared__ uint32_t s_w[32];
    const int t = threadIdx.x, lane = t & 31, w = t >> 5;
    const int64_t per = (m + 1023) / 1024, lo = (int64_t)t * per, hi = lo + per < m ? lo + per : m;
    uint32_t sum = 0;
    for (int64_t i = lo; i < hi; ++i) sum += hist[i];
    uint32_t inc = sum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const uint32_t x = __shfl_up_sync(FULL, inc, o); if (lane >= o) inc += x; }
    if (lane == 31) s_w[w] = inc;
    __syncthreads();
    if (w == 0) {
        uint32_t x = s_w[lane], y = x;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const uint32_t z = __shfl_up_sync(FULL, y, o); if (lane >= o) y += z; }
        s_w[lane] = y - x;                                           // exclusive over warps
    }
    __syncthreads();
    uint32_t run = s_w[w] + inc - sum;
    for (int64_t i = lo; i < hi; ++i) { const uint32_t c = hist[i]; hist[i] = run; run += c; }
}

__global__ void __launch_bounds__(RS_THREADS) sort_scatter_kernel(const uint64_t *__restrict__ kin, const uint32_t *__restrict__ vin, uint64_t *__restrict__ kout,
                                                                  uint32_t *__restrict__ vout, int64_t n, int shift, const uint32_t *__restrict__ hist, int n_tiles)
{
    __shared__ uint32_t s_cnt[RS_WARPS][256];
    const int t = threadIdx.x, w = t >> 5, lane = t & 31;
    for (int q = t; q < RS_WARPS * 256; q += RS_THREADS) (&s_cnt[0][0])[q] = 0;
    __syncthreads();
    const int64_t base = (int64_t)blockIdx.x * RS_TILE + w * (RS_TILE / RS_WARPS);       // this warp's consecutive rows
    uint64_t k[RS_ROUNDS]; uint32_t v[RS_ROUNDS], rk[RS_ROUNDS]; int d[RS_ROUNDS];
#pragma unroll
    for (int r = 0; r < RS_ROUNDS; ++r) {
        const int64_t i = base + r * 32 + lane;
        const bool ok = i < n;
        k[r] = ok ? kin[i] : 0; v[r] = ok ? vin[i] : 0;
        d[r] = ok ? (int)((k[r] >> shift) & 255u) : 256;
        const unsigned peers = __match_any_sync(FULL, d[r]);
        const unsigned lt = peers & ((1u << lane) - 1u);
        uint32_t before = 0;
        if (ok) before = s_cnt[w][d[r]];                             // rows of this warp's earlier rounds with the digit
        __syncwarp();
        if (ok && lt == 0) s_cnt[w][d[r]] = before + (uint32_t)__popc(peers);
        __syncwarp();
        rk[r] = before + (uint32_t)__popc(lt);
    }
    __syncthreads();
    {   // thread t <-> digit t: exclusive scan over the warps of the tile
        uint32_t acc = 0;
#pragma unroll
        for (int ww = 0; ww < RS_WARPS; ++ww) { const uint32_t c = s_cnt[ww][t]; s_cnt[ww][t] = acc; acc += c; }
    }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < RS_ROUNDS; ++r)
        if (d[r] < 256) {
            const uint32_t dst = hist[(size_t)d[r] * n_tiles + blockIdx.x] + s_cnt[w][d[r]] + rk[r];
            kout[dst] = k[r]; vout[dst] = v[r];
        }
}

__global__ void __launch_bounds__(RS_THREADS) rows_permute_kernel(DRows in, DRows out, const uint32_t *__restrict__ perm)
{
    const int64_t i = (int64_t)blockIdx.x * RS_THREADS + threadIdx.x;
    if (i >= in.n) return;
    const uint32_t s = perm[i];
    out.read_idx[i] = in.read_idx[s]; out.tid[i] = in.tid[s]; out.start[i] = in.start[s]; out.end[i] = in.end[s];
    out.is_rev[i] = in.is_rev[s]; out.ex_beg[i] = in.ex_beg[s]; out.ex_n[i] = in.ex_n[s];
}

int sort_tiles(int64_t n) { return (int)((n + RS_TILE - 1) / RS_TILE); }

void launch_sort_keys(const DRows &rows, const uint16_t *flag, uint64_t *keys, uint32_t *idx, unsigned long long *max_key, cudaStream_t st)
{
    if (rows.n <= 0) return;
    sort_keys_kernel<<<(unsigned)((rows.n + RS_THREADS - 1) / RS_THREADS), RS_THREADS, 0, st>>>(rows, flag, keys, idx, max_key); LRB_COUNT_LAUNCH();
}
// one pass on digit `shift / 8`: (kin, vin) -> (kout, vout); hist: 256 * sort_tiles(n) counters
void launch_sort_pass(const uint64_t *kin, const uint32_t *vin, uint64_t *kout, uint32_t *vout, int64_t n, int shift, uint32_t *hist, cudaStream_t st)
{
    if (n <= 0) return;
    const int nt = sort_tiles(n);
    sort_hist_kernel<<<nt, RS_THREADS, 0, st>>>(kin, n, shift, hist, nt); LRB_COUNT_LAUNCH();
    sort_scan_kernel<<<1, 1024, 0, st>>>(hist, (int64_t)256 * nt); LRB_COUNT_LAUNCH();
    sort_scatter_kernel<<<nt, RS_THREADS, 0, st>>>(kin, vin, kout, vout, n, shift, hist, nt); LRB_COUNT_LAUNCH();
}
void launch_rows_permute(const DRows &in, const DRows &out, const uint32_t *perm, cudaStream_t st)
{
    if (in.n <= 0) return;
    rows_permute_kernel<<<(unsigned)((in.n + RS_THREADS - 1) / RS_THREADS), RS_THREADS, 0, st>>>(in, out, perm); LRB_COUNT_LAUNCH();
}

}  // namespace lrbk
