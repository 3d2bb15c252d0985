// lrb_summary.cu -- K6: the "sets" of print_trans_summary() (update_gtf.c:421-587) as data-parallel passes.
//
// The reference builds five sets over updated_T by back-scans that test for an equal key first and stop at the first
// inserted entry with a smaller tid (add_simp_gene/_exon/_site/_sj, update_gtf.c:181-295).  For a tid-monotone stream
// that is "distinct by key, payload of the first occurrence, insertion order".  Split pieces carry trans_t.tid == 0
// (SURVEY Q14), which turns their gene / site / junction entries into barriers inside a chromosome block.  The exact
// rule used here (derivation in DESIGN.md):
//   * an element whose entry tid is 0 is inserted iff its key never occurred before among tid-0 elements
//     (genes: among all elements, gene equality ignores tid);
//   * an element with tid t > 0 is inserted iff no equal key occurred earlier within its segment, where a segment ends
//     at every *inserted* tid-0 element of the same set (genes: and it does not equal the gene of that barrier element
//     when the barrier lies inside the same chromosome block).
// Both are first-occurrence queries, answered with one lock-free open-addressing table keyed by (set, segment, tid,
// k1, k2) that keeps the minimum stream position per key; exon scores are atomically accumulated per key.
#include "lrb_common.cuh"
#include "lrb_kernels.cuh"
#include "lrb_summary.cuh"

namespace lrbk {

extern int64_t g_launches_summary;
int64_t g_launches_summary = 0;
#define LRB_COUNT_LAUNCH() (++g_launches_summary)

static constexpr uint64_t EMPTY = ~0ull;
static constexpr uint32_t SEG_TID0 = 0x1FFFFFFEu;   // pseudo segment of the tid-0 phase

LRB_DEVINL uint64_t mix64(uint64_t x) { x ^= x >> 33; x *= 0xFF51AFD7ED558CCDull; x ^= x >> 33; x *= 0xC4CEB9FE1A85EC53ull; x ^= x >> 33; return x; }
LRB_DEVINL uint64_t key_hi(int set, uint32_t seg, int tid) { return ((uint64_t)(((uint32_t)set << 29) | (seg & 0x1FFFFFFFu)) << 32) | (uint32_t)tid; }
LRB_DEVINL uint64_t key_lo(int k1, int k2) { return ((uint64_t)(uint32_t)k1 << 32) | (uint32_t)k2; }

// insert-or-find; returns the slot
LRB_DEVINL uint64_t tab_upsert(const HashTab &t, uint64_t hi, uint64_t lo)
{
    uint64_t s = __umul64hi(mix64(hi * 0x9E3779B97F4A7C15ull ^ mix64(lo)), t.cap);      // any capacity, not only powers of two
    for (;;) {
        unsigned long long p = atomicCAS((unsigned long long *)&t.slots[s].khi, (unsigned long long)EMPTY, (unsigned long long)hi);
        if (p == EMPTY || p == hi) {
            unsigned long long q = atomicCAS((unsigned long long *)&t.slots[s].klo, (unsigned long long)EMPTY, (unsigned long long)lo);
            if (q == EMPTY || q == lo) return s;
        }
        s = s + 1 == t.cap ? 0 : s + 1;
    }
}
LRB_DEVINL uint64_t tab_find(const HashTab &t, uint64_t hi, uint64_t lo)
{
    uint64_t s = __umul64hi(mix64(hi * 0x9E3779B97F4A7C15ull ^ mix64(lo)), t.cap);      // any capacity, not only powers of two
    for (;;) {
        const ulonglong2 k = *(const ulonglong2 *)&t.slots[s].khi;      // both key words in one 16-byte load
        const uint64_t a = k.x;
        if (a == hi && k.y == lo) return s;
        if (a == EMPTY) return EMPTY;
        s = s + 1 == t.cap ? 0 : s + 1;
    }
}
LRB_DEVINL void tab_min(const HashTab &t, uint64_t s, uint64_t pos) { atomicMin((unsigned long long *)&t.slots[s].minpos, (unsigned long long)pos); }

struct EntryView {
    int n; uint32_t gbeg; int fs, le; int t_tid, real_tid, rev, cov, gene, piece; uint32_t row;
};
LRB_DEVINL EntryView load_entry(const SummaryArgs &a, int64_t i)
{
    EntryView e; uint32_t c = a.upd.cand[i];
    e.row = a.list.row[c]; e.n = (int)a.list.cnt[c]; e.gbeg = a.rows.ex_beg[e.row] + a.list.lo[c];
    e.fs = a.upd.fs[i]; e.le = a.upd.le[i]; e.t_tid = a.upd.tid[i]; e.real_tid = a.rows.tid[e.row];
    e.rev = a.rows.is_rev[e.row]; e.cov = a.upd.cov[i]; e.piece = a.list.piece[c];
    int ref = a.ref[e.row];
    e.gene = ref >= 0 ? a.anno_gene[ref] : -1;
    return e;
}
LRB_DEVINL int ent_s(const SummaryArgs &a, const EntryView &e, int j) { return j == 0 ? e.fs : a.ex.es[e.gbeg + j]; }
LRB_DEVINL int ent_e(const SummaryArgs &a, const EntryView &e, int j) { return j == e.n - 1 ? e.le : a.ex.ee[e.gbeg + j]; }
LRB_DEVINL uint64_t pos_of(int64_t i, int j) { return ((uint64_t)i << 20) | (uint32_t)j; }

enum { SET_E = 0, SET_D = 1, SET_A = 2, SET_J = 3, SET_G = 4, SET_KG = 5, SET_PJ = 6, SET_PJ0 = 7 };

// Every pass below runs SG lanes per updated entry: lane l takes the exons l, l + SG, ... of the entry, so the dependent
// chain of table operations per thread is one or two elements long instead of the whole exon list (the passes are bound
// by the latency of those chains, not by bandwidth), and the lanes' counts are folded with group shuffles.
static constexpr int SG = 1;    // measured on B200: 8 lanes per entry is SLOWER (the passes are bound by the entry gathers and the random table sectors, not by the chain length)
#define SUM_ENTRY_PROLOGUE(n_limit)                                                                  \
    const int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) / SG;                         \
    const int gl = threadIdx.x % SG;                                                                 \
    if (i >= (n_limit)) return;                                                                      \
    const unsigned gm = group_mask<SG>();                                                            \
    (void)gm; (void)gl;

// element counts per set (sizes the table)
__global__ void sum_count_kernel(SummaryArgs a, unsigned long long *n_elems)
{
    const int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) / SG;
    const int gl = threadIdx.x % SG;
    int c = 0, cx = 0;                               // all elements; exon elements (one table key each -- the others can take two)
    if (i < (a.n_upd_dev ? (int64_t)*a.n_upd_dev : a.n_upd)) {
        EntryView e = load_entry(a, i);
        const uint8_t *f = a.ex.flag + e.gbeg;
        c = (gl == 0 && (a.sets & SUM_G)) ? 1 : 0;
        for (int j = gl; j < e.n && (a.sets & (SUM_E | SUM_DAJ)); j += SG) {
            uint8_t x = f[j];
            if (a.sets & SUM_E) cx += (x & LRB_F_NOVEL_EXON) != 0;
            if (j < e.n - 1 && (a.sets & SUM_DAJ)) c += ((x & LRB_F_NOVEL_DON) != 0) + ((x & LRB_F_NOVEL_ACC) != 0) + ((x & LRB_F_NOVEL_JUNC) != 0);
        }
        c += cx;
        if (a.probe && e.piece >= 0 && gl == 0) c += e.n;      // the junction keys of a split piece (cross-chromosome / cross-shard probe of phase 1 / 2)
    }
    for (int o = 16; o > 0; o >>= 1) { c += __shfl_xor_sync(FULL, c, o); cx += __shfl_xor_sync(FULL, cx, o); }
    if (lane_id() == 0 && c) { atomicAdd(n_elems, (unsigned long long)c); if (cx) atomicAdd(n_elems + 1, (unsigned long long)cx); }
}

// The exon set does not interact with the other four (plain first occurrence, its own keys), so it runs as its own chain
// of kernels on a side stream (insert -> count -> scan -> BED rows) next to the barrier-aware chain of the gene / site /
// junction sets: both are bound by dependent table look-ups, not by throughput, and overlap almost completely.
// `parts` bit 0: the exon elements are handled by this launch (0 when the exon chain runs separately).
__global__ void sum_exon_insert_kernel(SummaryArgs a)
{
    SUM_ENTRY_PROLOGUE(a.n_upd)
    EntryView e = load_entry(a, i);
    const uint8_t *f = a.ex.flag + e.gbeg;
    for (int j = gl; j < e.n; j += SG)
        if (f[j] & LRB_F_NOVEL_EXON) {
            uint64_t s = tab_upsert(a.tab, key_hi(SET_E, 0, e.real_tid), key_lo(ent_s(a, e, j), ent_e(a, e, j)));
            tab_min(a.tab, s, pos_of(i, j));
            atomicAdd(&a.tab.slots[s].score, e.cov);
        }
}
__global__ void sum_exon_count_kernel(SummaryArgs a)
{
    SUM_ENTRY_PROLOGUE(a.n_upd)
    EntryView e = load_entry(a, i);
    const uint8_t *f = a.ex.flag + e.gbeg;
    int ce = 0;
    for (int j = gl; j < e.n; j += SG)
        if (f[j] & LRB_F_NOVEL_EXON) {
            uint64_t s = tab_find(a.tab, key_hi(SET_E, 0, e.real_tid), key_lo(ent_s(a, e, j), ent_e(a, e, j)));
            ce += a.tab.slots[s].minpos == pos_of(i, j);
        }
    ce = group_sum<SG>(gm, ce);
    if (gl != 0) return;
    a.bed_cnt[i] = ce;
    if (ce) atomicAdd(&a.counts[SET_E], (uint32_t)ce);
}

LRB_DEVINL int shard_of(const int64_t *shard_end, int n_shards, int64_t i) { int k = 0; while (k < n_shards - 1 && i >= shard_end[k]) ++k; return k; }
// what must differ between a split piece and an entry with an equal junction for the pair to be a cross-locus meeting: the chromosome,
// or (gather root of a multi-GPU run, where the shards have settled the meetings inside themselves) the shard
LRB_DEVINL int xl_group(const SummaryArgs &a, const EntryView &e, int64_t i) { return a.n_shards > 0 ? shard_of(a.shard_end, a.n_shards, i) : e.real_tid; }

// phase 1: exons (all entries), tid-0 elements of D/A/J, every gene element (gene equality ignores tid)
__global__ void sum_phase1_kernel(SummaryArgs a, int parts)
{
    SUM_ENTRY_PROLOGUE(a.n_upd)
    EntryView e = load_entry(a, i);
    const uint8_t *f = a.ex.flag + e.gbeg;
    if (gl == 0 && (a.sets & SUM_G)) tab_min(a.tab, tab_upsert(a.tab, key_hi(SET_G, SEG_TID0, 0), key_lo(e.gene, 0)), pos_of(i, 0));
    if (e.piece >= 0 && gl == 0) {
        // a split piece scans the WHOLE of updated_T in the reference (its tid/start/end are 0, update_gtf.c:148 never stops it), so
        // it can meet an equal or partially matching chain on another chromosome.  Its junctions go into the table here and
        // phase 2 lets every entry probe them: a hit across chromosomes raises CNT_XLOCUS and the fold is replayed as one locus.
        atomicAdd(&a.counts[CNT_PARTIAL], 1u);       // partial-read transcripts
        const int grp = xl_group(a, e, i);
        if (a.probe && a.xs_pl && e.n > 1) { const uint32_t k = atomicAdd(&a.xs_cnt[0], 1u); if (k < a.xs_cap) a.xs_pl[k] = (uint32_t)i; else a.counts[CNT_XLOCUS] = 1u; }
        for (int j = 0; j < e.n - 1 && a.probe; ++j) {
            const uint64_t s = tab_upsert(a.tab, key_hi(SET_PJ, 0, 0), key_lo(ent_e(a, e, j), ent_s(a, e, j + 1)));
            atomicMin((unsigned long long *)&a.tab.slots[s].minpos, (unsigned long long)(uint32_t)grp); atomicMax(&a.tab.slots[s].pad, grp);
            if (j == 0) {
                const uint64_t s0 = tab_upsert(a.tab, key_hi(SET_PJ0, 0, 0), key_lo(ent_e(a, e, 0), ent_s(a, e, 1)));
                atomicMin((unsigned long long *)&a.tab.slots[s0].minpos, (unsigned long long)(uint32_t)grp); atomicMax(&a.tab.slots[s0].pad, grp);
            }
        }
    }
    const bool do_e = (parts & 1) && (a.sets & SUM_E), do_t0 = e.t_tid == 0 && (a.sets & SUM_DAJ);
    if (!do_e && !do_t0) return;
    for (int j = gl; j < e.n; j += SG) {
        uint8_t x = f[j];
        if (do_e && (x & LRB_F_NOVEL_EXON)) {
            uint64_t s = tab_upsert(a.tab, key_hi(SET_E, 0, e.real_tid), key_lo(ent_s(a, e, j), ent_e(a, e, j)));
            tab_min(a.tab, s, pos_of(i, j));
            atomicAdd(&a.tab.slots[s].score, e.cov);
        }
        if (do_t0 && j < e.n - 1) {
            if (x & LRB_F_NOVEL_DON) tab_min(a.tab, tab_upsert(a.tab, key_hi(SET_D, SEG_TID0, 0), key_lo(ent_e(a, e, j), 0)), pos_of(i, j));
            if (x & LRB_F_NOVEL_ACC) tab_min(a.tab, tab_upsert(a.tab, key_hi(SET_A, SEG_TID0, 0), key_lo(ent_s(a, e, j + 1), 0)), pos_of(i, j));
            if (x & LRB_F_NOVEL_JUNC) tab_min(a.tab, tab_upsert(a.tab, key_hi(SET_J, SEG_TID0, 0), key_lo(ent_e(a, e, j), ent_s(a, e, j + 1))), pos_of(i, j));
        }
    }
}

// phase 2: per entry, how many of its tid-0 elements were inserted (= barriers), per set; exon first occurrences
__global__ void sum_phase2_kernel(SummaryArgs a, int parts)
{
    SUM_ENTRY_PROLOGUE(a.n_upd)
    EntryView e = load_entry(a, i);
    const uint8_t *f = a.ex.flag + e.gbeg;
    int cd = 0, ca = 0, cj = 0, cg = 0, ce = 0;
    if (e.t_tid == 0 && gl == 0 && (a.sets & SUM_G)) {
        uint64_t s = tab_find(a.tab, key_hi(SET_G, SEG_TID0, 0), key_lo(e.gene, 0));
        cg = a.tab.slots[s].minpos == pos_of(i, 0);
    }
    if (a.probe && gl == 0 && e.n > 1 && a.counts[CNT_PARTIAL] != 0) {
        // cross-chromosome meeting of a split piece (see phase 1): the first junction of this entry among the junctions of a
        // piece, or a junction of this entry equal to the first junction of a piece -- the necessary condition of check_iden != -1
        bool hit = false;
        const int grp = xl_group(a, e, i);
        uint64_t s = tab_find(a.tab, key_hi(SET_PJ, 0, 0), key_lo(ent_e(a, e, 0), ent_s(a, e, 1)));
        if (s != EMPTY) hit = (int)(uint32_t)a.tab.slots[s].minpos != grp || a.tab.slots[s].pad != grp;
        for (int j = 0; j < e.n - 1 && !hit; ++j) {
            s = tab_find(a.tab, key_hi(SET_PJ0, 0, 0), key_lo(ent_e(a, e, j), ent_s(a, e, j + 1)));
            if (s != EMPTY) hit = (int)(uint32_t)a.tab.slots[s].minpos != grp || a.tab.slots[s].pad != grp;
        }
        if (hit) {
            // the necessary condition only; with the lists the pair is settled exactly by sum_xjoin_kernel (one shared junction is common
            // by chance in a deep data set, a chain the piece really merges into is not)
            if (a.xs_hx) { const uint32_t k = atomicAdd(&a.xs_cnt[1], 1u); if (k < a.xs_cap) a.xs_hx[k] = (uint32_t)i; else a.counts[CNT_XLOCUS] = 1u; }
            else a.counts[CNT_XLOCUS] = 1u;
        }
    }
    const bool do_e = (parts & 1) && (a.sets & SUM_E), do_t0 = e.t_tid == 0 && (a.sets & SUM_DAJ);
    for (int j = gl; j < e.n && (do_e || do_t0); j += SG) {
        uint8_t x = f[j];
        if (do_e && (x & LRB_F_NOVEL_EXON)) {
            uint64_t s = tab_find(a.tab, key_hi(SET_E, 0, e.real_tid), key_lo(ent_s(a, e, j), ent_e(a, e, j)));
            ce += a.tab.slots[s].minpos == pos_of(i, j);
        }
        if (do_t0 && j < e.n - 1) {
            if (x & LRB_F_NOVEL_DON) cd += a.tab.slots[tab_find(a.tab, key_hi(SET_D, SEG_TID0, 0), key_lo(ent_e(a, e, j), 0))].minpos == pos_of(i, j);
            if (x & LRB_F_NOVEL_ACC) ca += a.tab.slots[tab_find(a.tab, key_hi(SET_A, SEG_TID0, 0), key_lo(ent_s(a, e, j + 1), 0))].minpos == pos_of(i, j);
            if (x & LRB_F_NOVEL_JUNC) cj += a.tab.slots[tab_find(a.tab, key_hi(SET_J, SEG_TID0, 0), key_lo(ent_e(a, e, j), ent_s(a, e, j + 1)))].minpos == pos_of(i, j);
        }
    }
    ce = group_sum<SG>(gm, ce);
    if (e.t_tid == 0) { cd = group_sum<SG>(gm, cd); ca = group_sum<SG>(gm, ca); cj = group_sum<SG>(gm, cj); cg = group_sum<SG>(gm, cg); }
    if (gl != 0) return;
    a.bar_cnt[0 * a.n_upd + i] = cd; a.bar_cnt[1 * a.n_upd + i] = ca; a.bar_cnt[2 * a.n_upd + i] = cj; a.bar_cnt[3 * a.n_upd + i] = cg;
    if (parts & 1) a.bed_cnt[i] = ce;
    a.gene_bar[i] = cg ? (uint64_t)(i + 1) : 0;       // for the "last inserted tid-0 gene entry before x" max-scan
    if (cd | ca | cj | cg) {
        if (cd) atomicAdd(&a.counts[SET_D], (uint32_t)cd);
        if (ca) atomicAdd(&a.counts[SET_A], (uint32_t)ca);
        if (cj) atomicAdd(&a.counts[SET_J], (uint32_t)cj);
        if (cg) atomicAdd(&a.counts[SET_G], (uint32_t)cg);
    }
    if (ce) atomicAdd(&a.counts[SET_E], (uint32_t)ce);
}

// check_iden (gtf.c:54-92) of two entries on their internal boundaries, exact matching (-d 0), the end tests left out (a superset of the
// meetings under -D: the caller only grows more careful): 0 identical, 2 partial match, -1 unrelated
LRB_DEVINL int entry_chain_rel(const SummaryArgs &a, const EntryView &p, const EntryView &q)
{
    if (p.n < 2 || q.n < 2) return -1;
    if (p.n == q.n) {
        for (int i = 0; i < p.n - 1; ++i)
            if (a.ex.ee[p.gbeg + i] != a.ex.ee[q.gbeg + i] || a.ex.es[p.gbeg + i + 1] != a.ex.es[q.gbeg + i + 1]) return -1;
        return 0;
    }
    const EntryView &l = p.n > q.n ? p : q, &s = p.n > q.n ? q : p;
    const int s_e0 = a.ex.ee[s.gbeg], s_s1 = a.ex.es[s.gbeg + 1];
    for (int i = 0; i < l.n - 1; ++i)
        if (a.ex.ee[l.gbeg + i] == s_e0 && a.ex.es[l.gbeg + i + 1] == s_s1) {
            int j = 1;
            for (i = i + 1; i < l.n - 1 && j < s.n - 1; ++i, ++j)
                if (a.ex.ee[l.gbeg + i] != a.ex.ee[s.gbeg + j] || a.ex.es[l.gbeg + i + 1] != a.ex.es[s.gbeg + j + 1]) return -1;
            return 2;
        }
    return -1;
}
// pieces x hits: does an EARLIER entry of another chromosome / shard really absorb the piece?
__global__ void __launch_bounds__(256) sum_xjoin_kernel(SummaryArgs a)
{
    const uint32_t n_pl = min(a.xs_cnt[0], a.xs_cap), n_hx = min(a.xs_cnt[1], a.xs_cap);
    for (uint32_t h = blockIdx.y; h < n_hx; h += gridDim.y) {
        const int64_t xi = a.xs_hx[h];
        const EntryView x = load_entry(a, xi);
        const int gx = xl_group(a, x, xi);
        for (uint32_t k = blockIdx.x * blockDim.x + threadIdx.x; k < n_pl; k += gridDim.x * blockDim.x) {
            const int64_t pi = a.xs_pl[k];
            if (xi >= pi) continue;
            const EntryView p = load_entry(a, pi);
            if (xl_group(a, p, pi) == gx || p.real_tid == x.real_tid) continue;      // the same chromosome is the fold's own business
            if (entry_chain_rel(a, p, x) >= 0) a.counts[CNT_XLOCUS] = 1u;
        }
    }
}

// phase 3: tid>0 elements go into their segment
__global__ void sum_phase3_kernel(SummaryArgs a)
{
    SUM_ENTRY_PROLOGUE(a.n_upd)
    EntryView e = load_entry(a, i);
    if (e.t_tid == 0) return;
    const uint8_t *f = a.ex.flag + e.gbeg;
    const uint32_t sd = a.bar_seg[0 * a.n_upd + i], sa = a.bar_seg[1 * a.n_upd + i], sj = a.bar_seg[2 * a.n_upd + i], sg = a.bar_seg[3 * a.n_upd + i];
    if (gl == 0 && (a.sets & SUM_G)) tab_min(a.tab, tab_upsert(a.tab, key_hi(SET_G, sg, e.t_tid), key_lo(e.gene, 0)), pos_of(i, 0));
    if (!(a.sets & SUM_DAJ)) return;
    for (int j = gl; j < e.n - 1; j += SG) {
        uint8_t x = f[j];
        if (x & LRB_F_NOVEL_DON) tab_min(a.tab, tab_upsert(a.tab, key_hi(SET_D, sd, e.t_tid), key_lo(ent_e(a, e, j), 0)), pos_of(i, j));
        if (x & LRB_F_NOVEL_ACC) tab_min(a.tab, tab_upsert(a.tab, key_hi(SET_A, sa, e.t_tid), key_lo(ent_s(a, e, j + 1), 0)), pos_of(i, j));
        if (x & LRB_F_NOVEL_JUNC) tab_min(a.tab, tab_upsert(a.tab, key_hi(SET_J, sj, e.t_tid), key_lo(ent_e(a, e, j), ent_s(a, e, j + 1))), pos_of(i, j));
    }
}
__global__ void sum_phase4_kernel(SummaryArgs a)
{
    SUM_ENTRY_PROLOGUE(a.n_upd)
    EntryView e = load_entry(a, i);
    if (e.t_tid == 0) return;
    const uint8_t *f = a.ex.flag + e.gbeg;
    const uint32_t sd = a.bar_seg[0 * a.n_upd + i], sa = a.bar_seg[1 * a.n_upd + i], sj = a.bar_seg[2 * a.n_upd + i], sg = a.bar_seg[3 * a.n_upd + i];
    int cd = 0, ca = 0, cj = 0, cg = 0;
    if (gl == 0 && (a.sets & SUM_G)) {
        uint64_t s = tab_find(a.tab, key_hi(SET_G, sg, e.t_tid), key_lo(e.gene, 0));
        bool first = a.tab.slots[s].minpos == pos_of(i, 0);
        uint64_t bar = a.gene_bar[i];                // index+1 of the last inserted tid-0 gene entry before i (inclusive scan, own value 0)
        if (first && bar) {
            EntryView b = load_entry(a, (int64_t)bar - 1);
            if (b.real_tid == e.t_tid && b.gene == e.gene) first = false;     // equal to the barrier itself (match precedes stop)
        }
        cg = first;
    }
    for (int j = gl; j < e.n - 1 && (a.sets & SUM_DAJ); j += SG) {
        uint8_t x = f[j];
        if (x & LRB_F_NOVEL_DON) cd += a.tab.slots[tab_find(a.tab, key_hi(SET_D, sd, e.t_tid), key_lo(ent_e(a, e, j), 0))].minpos == pos_of(i, j);
        if (x & LRB_F_NOVEL_ACC) ca += a.tab.slots[tab_find(a.tab, key_hi(SET_A, sa, e.t_tid), key_lo(ent_s(a, e, j + 1), 0))].minpos == pos_of(i, j);
        if (x & LRB_F_NOVEL_JUNC) cj += a.tab.slots[tab_find(a.tab, key_hi(SET_J, sj, e.t_tid), key_lo(ent_e(a, e, j), ent_s(a, e, j + 1)))].minpos == pos_of(i, j);
    }
    cd = group_sum<SG>(gm, cd); ca = group_sum<SG>(gm, ca); cj = group_sum<SG>(gm, cj);
    if (gl != 0) return;
    if (cd) atomicAdd(&a.counts[SET_D], (uint32_t)cd);
    if (ca) atomicAdd(&a.counts[SET_A], (uint32_t)ca);
    if (cj) atomicAdd(&a.counts[SET_J], (uint32_t)cj);
    if (cg) atomicAdd(&a.counts[SET_G], (uint32_t)cg);
}

// BED rows: first occurrences of the exon set in stream order (bed_off = exclusive scan of bed_cnt); the lanes of an
// entry rank their rows by exon index with a ballot per round of SG exons
__global__ void sum_bed_kernel(SummaryArgs a)
{
    SUM_ENTRY_PROLOGUE(a.n_upd)
    if (a.bed_cnt[i] == 0) return;
    EntryView e = load_entry(a, i);
    const uint8_t *f = a.ex.flag + e.gbeg;
    uint32_t o = a.bed_off[i];
    const int sh = (lane_id() / SG) * SG;
    for (int j0 = 0; j0 < e.n; j0 += SG) {
        const int j = j0 + gl;
        bool first = false; int s0 = 0, e0 = 0; uint64_t s = 0;
        if (j < e.n && (f[j] & LRB_F_NOVEL_EXON)) {
            s0 = ent_s(a, e, j); e0 = ent_e(a, e, j);
            s = tab_find(a.tab, key_hi(SET_E, 0, e.real_tid), key_lo(s0, e0));
            first = a.tab.slots[s].minpos == pos_of(i, j);
        }
        const unsigned b = (__ballot_sync(gm, first) >> sh) & lane_bits<SG>();
        if (first) {
            const uint32_t k = o + __popc(b & ((1u << gl) - 1u));
            a.bed_tid[k] = e.real_tid; a.bed_start[k] = s0; a.bed_end[k] = e0; a.bed_score[k] = a.tab.slots[s].score + 1;        // the table is filled with 0xFF: scores start at -1
            a.bed_type[k] = e.n > 1 ? ((j == 0 || j == e.n - 1) ? 0 : 1) : 2; a.bed_rev[k] = (uint8_t)e.rev;
        }
        o += __popc(b);
    }
}

// the six scans between phase 2 and phase 3 in one launch: blockIdx.y selects the sequence (0-3 exclusive sums of the
// barrier counts, 4 inclusive running max of gene_bar in place, 5 exclusive sum of bed_cnt with its total); every
// sequence has its own ticket and look-back words
static constexpr int MS_THREADS = 256, MS_ITEMS = 8;
__global__ void __launch_bounds__(MS_THREADS) sum_scans_kernel(SummaryArgs a, uint64_t *tile_state, uint32_t *tickets, int n_tiles, uint64_t *bed_total, int which0)
{
    __shared__ uint32_t s_scan[33]; __shared__ uint64_t s_w[MS_THREADS / 32]; __shared__ uint32_t s_tile; __shared__ uint64_t s_excl;
    const int which = which0 + blockIdx.y;
    if (threadIdx.x == 0) s_tile = atomicAdd(tickets + which, 1u);
    __syncthreads();
    const int tile = (int)s_tile, lane = lane_id(), w = warp_id();
    uint64_t *state = tile_state + (size_t)which * n_tiles;
    const int64_t n = a.n_upd, base = ((int64_t)tile * MS_THREADS + threadIdx.x) * MS_ITEMS;
    if (which == 4) {
        uint64_t *data = a.gene_bar;
        uint64_t v[MS_ITEMS], run = 0;
#pragma unroll
        for (int i = 0; i < MS_ITEMS; ++i) { v[i] = (base + i < n) ? data[base + i] : 0; run = v[i] > run ? v[i] : run; v[i] = run; }
        uint64_t inc = run;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { uint64_t t = __shfl_up_sync(FULL, inc, o); if (lane >= o && t > inc) inc = t; }
        if (lane == 31) s_w[w] = inc;
        __syncthreads();
        uint64_t pre = 0, tot = 0;
        for (int k = 0; k < MS_THREADS / 32; ++k) { const uint64_t x = s_w[k]; if (k < w && x > pre) pre = x; if (x > tot) tot = x; }
        uint64_t left = __shfl_up_sync(FULL, inc, 1); if (lane == 0) left = 0;
        const uint64_t tpre = left > pre ? left : pre;
        if (w == 0) { uint64_t e = lookback_exclusive(state, tile, tot, OpMax()); if (lane == 0) s_excl = e; }
        __syncthreads();
        const uint64_t ex = s_excl > tpre ? s_excl : tpre;
#pragma unroll
        for (int i = 0; i < MS_ITEMS; ++i) if (base + i < n) data[base + i] = v[i] > ex ? v[i] : ex;
        return;
    }
    const uint32_t *in = which == 5 ? a.bed_cnt : a.bar_cnt + (int64_t)which * n;
    uint32_t *out = which == 5 ? a.bed_off : a.bar_seg + (int64_t)which * n;
    uint32_t v[MS_ITEMS], sum = 0;
#pragma unroll
    for (int i = 0; i < MS_ITEMS; ++i) { v[i] = (base + i < n) ? in[base + i] : 0; sum += v[i]; }
    uint32_t tot, excl = block_excl_sum(sum, s_scan, &tot);
    if (w == 0) { uint64_t e = lookback_exclusive(state, tile, tot, OpAdd()); if (lane == 0) s_excl = e; }
    __syncthreads();
    uint32_t o = (uint32_t)s_excl + excl;
#pragma unroll
    for (int i = 0; i < MS_ITEMS; ++i) if (base + i < n) { out[base + i] = o; o += v[i]; }
    if (which == 5 && tile == n_tiles - 1 && threadIdx.x == 0) *bed_total = s_excl + tot;
}

// genes of the known reads (update_gtf.c:503-506): distinct (tid, gene) over bam_T rows flagged known
__global__ void sum_known_genes_kernel(SummaryArgs a, const uint32_t *__restrict__ cls, int64_t n_rows, int pass)
{
    int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n_rows || !(cls[r] & LRB_C_KNOWN)) return;
    int ref = a.ref[r]; int gene = ref >= 0 ? a.anno_gene[ref] : -1;
    uint64_t hi = key_hi(SET_KG, 0, a.rows.tid[r]), lo = key_lo(gene, 0);
    if (pass == 0) tab_min(a.tab, tab_upsert(a.tab, hi, lo), (uint64_t)r);
    else if (a.tab.slots[tab_find(a.tab, hi, lo)].minpos == (uint64_t)r) {
        const uint32_t k = atomicAdd(&a.counts[SET_KG], 1u);
        if (a.kg_pairs) a.kg_pairs[k] = make_int2(a.rows.tid[r], gene);       // for the union across shards (order is irrelevant)
    }
}

// ---- helpers of the gather root (lrb_multi.cu)
__global__ void table_view_kernel(GatheredTable t, uint32_t *ident, uint32_t *zeros, uint32_t *cnt, int32_t *fs, int32_t *le)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= t.n) return;
    const uint32_t lo = t.exon_off[i], hi = t.exon_off[i + 1];
    ident[i] = (uint32_t)i; zeros[i] = 0; cnt[i] = hi - lo; fs[i] = t.es[lo]; le[i] = t.ee[hi - 1];
}
void launch_table_view(const GatheredTable &t, uint32_t *ident, uint32_t *zeros, uint32_t *cnt, int32_t *fs, int32_t *le, cudaStream_t st)
{
    if (t.n <= 0) return;
    table_view_kernel<<<(unsigned)((t.n + 255) / 256), 256, 0, st>>>(t, ident, zeros, cnt, fs, le); LRB_COUNT_LAUNCH();
}

// pass 0: every tid-0 site / junction element leaves the lowest and the highest shard that holds its key; pass 1: a key seen in two shards
__global__ void tid0_coincidence_kernel(SummaryArgs a, int pass, uint32_t *flag)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= a.n_upd) return;
    EntryView e = load_entry(a, i);
    if (e.t_tid != 0) return;
    const int sh = shard_of(a.shard_end, a.n_shards, i);
    const uint8_t *f = a.ex.flag + e.gbeg;
    for (int j = 0; j < e.n - 1; ++j) {
        const uint8_t x = f[j];
        for (int q = 0; q < 3; ++q) {
            if (!(x & (LRB_F_NOVEL_DON << q))) continue;
            const uint64_t hi = key_hi(SET_D + q, SEG_TID0, 0);
            const uint64_t lo = q == 0 ? key_lo(ent_e(a, e, j), 0) : q == 1 ? key_lo(ent_s(a, e, j + 1), 0) : key_lo(ent_e(a, e, j), ent_s(a, e, j + 1));
            if (pass == 0) { const uint64_t s = tab_upsert(a.tab, hi, lo); atomicMin((unsigned long long *)&a.tab.slots[s].minpos, (unsigned long long)sh); atomicMax(&a.tab.slots[s].pad, sh); }
            else { const uint64_t s = tab_find(a.tab, hi, lo); if ((int)a.tab.slots[s].minpos != a.tab.slots[s].pad) *flag = 1u; }
        }
    }
}
void launch_tid0_coincidence(const SummaryArgs &a, uint32_t *flag, cudaStream_t st)
{
    if (a.n_upd <= 0 || a.n_shards <= 1) return;
    for (int pass = 0; pass < 2; ++pass) { tid0_coincidence_kernel<<<(unsigned)((a.n_upd + 255) / 256), 256, 0, st>>>(a, pass, flag); LRB_COUNT_LAUNCH(); }
}

__global__ void pairs_distinct_kernel(HashTab tab, const int2 *__restrict__ pairs, int64_t n, int pass, uint32_t *out)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint64_t hi = key_hi(SET_KG, 0, pairs[i].x), lo = key_lo(pairs[i].y, 0);
    if (pass == 0) tab_min(tab, tab_upsert(tab, hi, lo), (uint64_t)i);
    else if (tab.slots[tab_find(tab, hi, lo)].minpos == (uint64_t)i) atomicAdd(out, 1u);
}
void launch_pairs_distinct(const HashTab &tab, const int2 *pairs, int64_t n, uint32_t *out, cudaStream_t st)
{
    if (n <= 0) return;
    for (int pass = 0; pass < 2; ++pass) { pairs_distinct_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(tab, pairs, n, pass, out); LRB_COUNT_LAUNCH(); }
}

static inline unsigned nblk(int64_t n) { return (unsigned)((n + 255) / 256); }
static inline unsigned nblk_g(int64_t n) { return (unsigned)((n * SG + 255) / 256); }       // SG lanes per entry

void launch_summary_count(const SummaryArgs &a, unsigned long long *n_elems, cudaStream_t st)
{
    if (a.n_upd <= 0) return;
    sum_count_kernel<<<nblk_g(a.n_upd), 256, 0, st>>>(a, n_elems); LRB_COUNT_LAUNCH();
}
// tile_state: 6 * (n_upd / 2048 + 1) words; tickets: 6 words (zeroed here).  st_exon != st: the exon chain (and the BED rows)
// runs on st_exon between ev_fork and ev_join; st waits for ev_join at the end
void launch_summary_sets(const SummaryArgs &a, const uint32_t *cls, int64_t n_rows, uint64_t *tile_state, uint32_t *tickets, uint64_t *bed_total,
                         cudaStream_t st, cudaStream_t st_exon, cudaEvent_t ev_fork, cudaEvent_t ev_join)
{
    const bool split = st_exon != st && a.n_upd > 0 && (a.sets & SUM_E);
    const int n_tiles = (int)((a.n_upd + MS_THREADS * MS_ITEMS - 1) / (MS_THREADS * MS_ITEMS));
    if (a.n_upd > 0) { cudaMemsetAsync(tile_state, 0, (size_t)n_tiles * 6 * 8, st); cudaMemsetAsync(tickets, 0, 6 * 4, st); }
    if (split) {
        cudaEventRecord(ev_fork, st); cudaStreamWaitEvent(st_exon, ev_fork, 0);
        sum_exon_insert_kernel<<<nblk_g(a.n_upd), 256, 0, st_exon>>>(a); LRB_COUNT_LAUNCH();
        sum_exon_count_kernel<<<nblk_g(a.n_upd), 256, 0, st_exon>>>(a); LRB_COUNT_LAUNCH();
        sum_scans_kernel<<<dim3((unsigned)n_tiles, 1), MS_THREADS, 0, st_exon>>>(a, tile_state, tickets, n_tiles, bed_total, 5); LRB_COUNT_LAUNCH();
        sum_bed_kernel<<<nblk_g(a.n_upd), 256, 0, st_exon>>>(a); LRB_COUNT_LAUNCH();
        cudaEventRecord(ev_join, st_exon);
    }
    const int parts = split ? 0 : 1;
    const bool kg = n_rows > 0 && (a.sets & SUM_KG), segs = (a.sets & (SUM_G | SUM_DAJ)) != 0;
    if (kg) { sum_known_genes_kernel<<<nblk(n_rows), 256, 0, st>>>(a, cls, n_rows, 0); LRB_COUNT_LAUNCH(); }
    if (a.n_upd > 0) {
        sum_phase1_kernel<<<nblk_g(a.n_upd), 256, 0, st>>>(a, parts); LRB_COUNT_LAUNCH();
        sum_phase2_kernel<<<nblk_g(a.n_upd), 256, 0, st>>>(a, parts); LRB_COUNT_LAUNCH();
        if (a.probe && a.xs_pl && a.xs_hx) { sum_xjoin_kernel<<<dim3(64, 64), 256, 0, st>>>(a); LRB_COUNT_LAUNCH(); }
        if (segs || (!split && (a.sets & SUM_E))) { sum_scans_kernel<<<dim3((unsigned)n_tiles, split ? 5 : 6), MS_THREADS, 0, st>>>(a, tile_state, tickets, n_tiles, bed_total, 0); LRB_COUNT_LAUNCH(); }
        if (segs) {
            sum_phase3_kernel<<<nblk_g(a.n_upd), 256, 0, st>>>(a); LRB_COUNT_LAUNCH();
            sum_phase4_kernel<<<nblk_g(a.n_upd), 256, 0, st>>>(a); LRB_COUNT_LAUNCH();
        }
    }
    if (a.n_upd <= 0 || !(a.sets & SUM_E)) cudaMemsetAsync(bed_total, 0, 8, st);
    if (kg) { sum_known_genes_kernel<<<nblk(n_rows), 256, 0, st>>>(a, cls, n_rows, 1); LRB_COUNT_LAUNCH(); }
    if (split) cudaStreamWaitEvent(st, ev_join, 0);
    else if (a.n_upd > 0 && (a.sets & SUM_E)) { sum_bed_kernel<<<nblk_g(a.n_upd), 256, 0, st>>>(a); LRB_COUNT_LAUNCH(); }
}

}  // namespace lrbk
