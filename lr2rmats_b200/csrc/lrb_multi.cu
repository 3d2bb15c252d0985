// lrb_multi.cu -- the multi-GPU layer of the C ABI: one process per GPU, one NCCL communicator per context.
//
//   lrb_tables_broadcast   the annotation / remove / SJ tables go from the root's HBM to every rank's HBM (ncclBroadcast over
//                          NVLink); no rank but the root parses or uploads them
//   lrb_update_gather      after lrb_update_run on every rank (each on its locus-aligned shard of ONE sorted read stream):
//                          counts all-gathered, the per-shard updated_T tables / BED rows / known-gene pairs sent to rank 0
//                          (ncclSend / ncclRecv, a gatherv), and the canonical merge on rank 0:
//                            * tables and BED rows concatenate in shard order;
//                            * counters add up, EXCEPT the gene sets -- a gene can span two loci, and gene equality ignores the
//                              chromosome (update_gtf.c:176-179): Updated_Genes is recomputed over the gathered table with the
//                              same barrier-aware kernels as on one GPU, Genes_of_Known... is the union of the shards' pairs;
//                            * the tid-0 site / junction keys of split pieces are probed across shards and the
//                              three sets recomputed over the gathered table when two shards share one;
//                            * a split piece whose junction also exists in ANOTHER shard could have been absorbed there in the
//                              reference (Q14.2): that is reported as LRB_E_XSHARD, the caller reruns unsharded.
//
// NCCL is bound at run time (dlopen of libnccl.so.2): the library loads on a box without NCCL, and inside a process that already
// carries NCCL (torch) it binds to that copy instead of a second one.
#include <dlfcn.h>
#include <nccl.h>
#include <vector>
#include "lrb_ctx.cuh"

namespace lrbk {

struct NcclApi {
    void *h = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*Broadcast)(const void *, void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllGather)(const void *, void *, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Send)(const void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char *(*GetErrorString)(ncclResult_t) = nullptr;
};

static NcclApi *nccl_api(std::string *err)
{
    static NcclApi api; static bool tried = false; static std::string why;
    if (!tried) {
        tried = true;
        const char *names[] = {getenv("LRB_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
        for (const char *nm : names) { if (nm && *nm && (api.h = dlopen(nm, RTLD_NOW | RTLD_LOCAL))) break; }
        if (!api.h) why = std::string("libnccl.so.2 not found (") + (dlerror() ? dlerror() : "dlopen failed") + "); set LRB_NCCL_LIB";
        else {
#define LRB_SYM(field, name) do { *(void **)&api.field = dlsym(api.h, name); if (!api.field) why = std::string("NCCL symbol missing: ") + name; } while (0)
            LRB_SYM(GetUniqueId, "ncclGetUniqueId"); LRB_SYM(CommInitRank, "ncclCommInitRank"); LRB_SYM(CommDestroy, "ncclCommDestroy");
            LRB_SYM(Broadcast, "ncclBroadcast"); LRB_SYM(AllGather, "ncclAllGather"); LRB_SYM(Send, "ncclSend"); LRB_SYM(Recv, "ncclRecv");
            LRB_SYM(GroupStart, "ncclGroupStart"); LRB_SYM(GroupEnd, "ncclGroupEnd"); LRB_SYM(GetErrorString, "ncclGetErrorString");
#undef LRB_SYM
        }
    }
    if (!why.empty()) { if (err) *err = why; return nullptr; }
    return &api;
}

static constexpr int META_WORDS = 40, MAX_RANKS = 64;
enum { M_NU = 0, M_NE, M_NBED, M_NKG, M_NAME_BASE, M_XLOCUS, M_PARTIAL, M_WANT_SUMMARY, M_SUMMARY0 /* .. + LRB_S_COUNT */ };

struct MultiState {
    NcclApi *api = nullptr; ncclComm_t comm = nullptr; int rank = 0, n_ranks = 1;
    Buf meta_dev, shard_end_dev, flags_dev; PBuf meta_host;
    // gathered on the root
    Buf g_name, g_piece, g_ttid, g_tstart, g_tend, g_trev, g_etid, g_erev, g_cov, g_ref, g_off, g_es, g_ee, g_flag;
    Buf g_bd_tid, g_bd_s, g_bd_e, g_bd_sc, g_bd_ty, g_bd_rv, g_kg;
    Buf v_ident, v_zeros, v_cnt, v_fs, v_le;          // the gathered table in the shape the set kernels read
    Buf tab, y_barcnt, y_barseg, y_genebar, y_bedcnt, y_bedoff, y_counts, y_nelem, tiles, xs_pl, xs_hx, xs_cnt;
    int64_t n = 0, n_exon = 0, n_bed = 0; int32_t summary[LRB_S_COUNT]; bool have = false, daj_recomputed = false;
    PBuf p[24];
    float ms_gather = 0, ms_merge = 0; cudaEvent_t ev[3] = {nullptr, nullptr, nullptr};
};

void multi_release(lrb_ctx *c)
{
    MultiState *m = c->multi;
    if (!m) return;
    if (m->comm && m->api) m->api->CommDestroy(m->comm);
    Buf *bufs[] = {&m->meta_dev, &m->shard_end_dev, &m->flags_dev, &m->g_name, &m->g_piece, &m->g_ttid, &m->g_tstart, &m->g_tend, &m->g_trev, &m->g_etid, &m->g_erev,
                   &m->g_cov, &m->g_ref, &m->g_off, &m->g_es, &m->g_ee, &m->g_flag, &m->g_bd_tid, &m->g_bd_s, &m->g_bd_e, &m->g_bd_sc, &m->g_bd_ty, &m->g_bd_rv, &m->g_kg,
                   &m->v_ident, &m->v_zeros, &m->v_cnt, &m->v_fs, &m->v_le, &m->tab, &m->y_barcnt, &m->y_barseg, &m->y_genebar, &m->y_bedcnt, &m->y_bedoff, &m->y_counts,
                   &m->y_nelem, &m->tiles, &m->xs_pl, &m->xs_hx, &m->xs_cnt};
    for (Buf *b : bufs) b->release();
    m->meta_host.release();
    for (PBuf &p : m->p) p.release();
    for (cudaEvent_t &e : m->ev) if (e) cudaEventDestroy(e);
    delete m; c->multi = nullptr;
}

#define NK(call)                                                                                                           \
    do { ncclResult_t r_ = (call); if (r_ != ncclSuccess) return fail(c, LRB_E_NCCL, std::string(#call) + ": " + m->api->GetErrorString(r_)); } while (0)

// exon offsets of shard r start at 0: shift them to the shard's place in the gathered pools; name_idx becomes an index into
// the whole read stream
struct RebaseArgs { int64_t row_end[MAX_RANKS]; int64_t exon_base[MAX_RANKS]; int64_t name_base[MAX_RANKS]; int n; };
__global__ void rebase_kernel(RebaseArgs a, uint32_t *exon_off, uint32_t *name_idx, int64_t n, int64_t n_exon)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i == 0) exon_off[n] = (uint32_t)n_exon;
    if (i >= n) return;
    int k = 0; while (k < a.n - 1 && i >= a.row_end[k]) ++k;
    exon_off[i] += (uint32_t)a.exon_base[k]; name_idx[i] += (uint32_t)a.name_base[k];
}

}  // namespace lrbk

using namespace lrbk;

extern "C" {

int lrb_comm_id(void *id_out)
{
    if (!id_out) return LRB_E_ARG;
    std::string why; NcclApi *api = nccl_api(&why);
    if (!api) return LRB_E_NCCL;
    ncclUniqueId id;
    if (api->GetUniqueId(&id) != ncclSuccess) return LRB_E_NCCL;
    static_assert(sizeof(ncclUniqueId) == LRB_COMM_ID_BYTES, "ncclUniqueId size");
    memcpy(id_out, &id, sizeof id);
    return LRB_OK;
}

int lrb_comm_init(lrb_ctx *c, const void *id_bytes, int rank, int n_ranks)
{
    if (!c || !id_bytes || n_ranks < 1 || n_ranks > MAX_RANKS || rank < 0 || rank >= n_ranks) return LRB_E_ARG;
    CK(cudaSetDevice(c->device));
    std::string why; NcclApi *api = nccl_api(&why);
    if (!api) return fail(c, LRB_E_NCCL, why);
    multi_release(c);
    MultiState *m = new MultiState(); c->multi = m;
    m->api = api; m->rank = rank; m->n_ranks = n_ranks;
    ncclUniqueId id; memcpy(&id, id_bytes, sizeof id);
    NK(api->CommInitRank(&m->comm, n_ranks, id, rank));
    NEED(m->meta_dev, (size_t)MAX_RANKS * META_WORDS * 8); NEED(m->shard_end_dev, MAX_RANKS * 8); NEED(m->flags_dev, 64);
    NEEDP(m->meta_host, (size_t)MAX_RANKS * META_WORDS * 8);
    for (cudaEvent_t &e : m->ev) CK(cudaEventCreate(&e));
    c->want_kg_pairs = true;
    return LRB_OK;
}

int lrb_comm_destroy(lrb_ctx *c)
{
    if (!c) return LRB_E_ARG;
    cudaSetDevice(c->device); cudaStreamSynchronize(c->st);
    multi_release(c); c->want_kg_pairs = false;
    return LRB_OK;
}

int lrb_comm_rank(const lrb_ctx *c, int *rank, int *n_ranks)
{
    if (!c || !c->multi) return LRB_E_ARG;
    if (rank) *rank = c->multi->rank;
    if (n_ranks) *n_ranks = c->multi->n_ranks;
    return LRB_OK;
}

// ------------------------------------------------------------------------------------------------ tables over NVLink
int lrb_tables_broadcast(lrb_ctx *c, int root, const lrb_anno *anno, const lrb_anno *rm, const lrb_sj *sj)
{
    if (!c) return LRB_E_ARG;
    MultiState *m = c->multi;
    if (!m) return fail(c, LRB_E_ARG, "lrb_tables_broadcast: no communicator (lrb_comm_init)");
    if (root < 0 || root >= m->n_ranks) return LRB_E_ARG;
    CK(cudaSetDevice(c->device));
    int rc;
    const bool is_root = m->rank == root;
    if (is_root) {                                   // the root builds the device tables (prefix-max keys, remove index) once
        if ((rc = lrb_anno_upload(c, anno))) return rc;
        if ((rc = lrb_rm_upload(c, rm))) return rc;
        if ((rc = lrb_sj_upload(c, sj))) return rc;
    }
    int64_t *hd = m->meta_host.as<int64_t>();
    if (is_root) { hd[0] = c->anno.n; hd[1] = c->anno.n_exon; hd[2] = c->sj.n; hd[3] = c->rm.n_groups; hd[4] = c->rm.n; }
    if (is_root) CK(cudaMemcpyAsync(m->meta_dev.p, hd, 5 * 8, cudaMemcpyHostToDevice, c->st));
    NK(m->api->Broadcast(m->meta_dev.p, m->meta_dev.p, 5, ncclInt64, root, m->comm, c->st));
    CK(cudaMemcpyAsync(hd, m->meta_dev.p, 5 * 8, cudaMemcpyDeviceToHost, c->st));
    CK(cudaStreamSynchronize(c->st));
    const int64_t na = hd[0], nae = hd[1], ns = hd[2], ng = hd[3], nr = hd[4];
    struct Item { Buf *b; size_t bytes; };
    std::vector<Item> items;
    if (na) {
        items = {{&c->a_tid, (size_t)na * 4}, {&c->a_start, (size_t)na * 4}, {&c->a_end, (size_t)na * 4}, {&c->a_gene, (size_t)na * 4}, {&c->a_rev, (size_t)na},
                 {&c->a_off, ((size_t)na + 1) * 4}, {&c->a_es, (size_t)nae * 4}, {&c->a_ee, (size_t)nae * 4}, {&c->a_pmax, (size_t)na * 8}, {&c->a_mono, (size_t)na}};
    }
    if (ns) for (Item it : {Item{&c->s_tid, (size_t)ns * 4}, Item{&c->s_don, (size_t)ns * 4}, Item{&c->s_acc, (size_t)ns * 4}, Item{&c->s_u, (size_t)ns * 4},
                            Item{&c->s_m, (size_t)ns * 4}, Item{&c->s_pmax, (size_t)ns * 8}, Item{&c->s_dkey, (size_t)ns * 8}}) items.push_back(it);
    if (nr) for (Item it : {Item{&c->r_gtid, (size_t)ng * 4}, Item{&c->r_goff, ((size_t)ng + 1) * 4}, Item{&c->r_start, (size_t)nr * 4}, Item{&c->r_pmax, (size_t)nr * 4}}) items.push_back(it);
    if (!is_root) for (Item &it : items) NEED(*it.b, std::max<size_t>(it.bytes, 1));
    NK(m->api->GroupStart());
    for (Item &it : items) if (it.bytes) NK(m->api->Broadcast(it.b->p, it.b->p, it.bytes, ncclChar, root, m->comm, c->st));
    NK(m->api->GroupEnd());
    if (!is_root) {
        c->anno = DAnno{}; c->sj = DSj{}; c->rm = DRmIndex{};
        if (na) {
            c->anno.n = (int32_t)na; c->anno.n_exon = nae; c->anno.tid = c->a_tid.as<int32_t>(); c->anno.start = c->a_start.as<int32_t>();
            c->anno.end = c->a_end.as<int32_t>(); c->anno.gene = c->a_gene.as<int32_t>(); c->anno.is_rev = c->a_rev.as<uint8_t>();
            c->anno.exon_off = c->a_off.as<uint32_t>(); c->anno.es = c->a_es.as<int32_t>(); c->anno.ee = c->a_ee.as<int32_t>();
            c->anno.pmax_key = c->a_pmax.as<uint64_t>(); c->anno.mono = c->a_mono.as<uint8_t>();
        }
        if (ns) {
            c->sj.n = ns; c->sj.tid = c->s_tid.as<int32_t>(); c->sj.don = c->s_don.as<int32_t>(); c->sj.acc = c->s_acc.as<int32_t>();
            c->sj.cnt_u = c->s_u.as<int32_t>(); c->sj.cnt_m = c->s_m.as<int32_t>(); c->sj.pmax_key = c->s_pmax.as<uint64_t>(); c->sj.don_key = c->s_dkey.as<uint64_t>();
        }
        if (nr) {
            c->rm.n_groups = (int32_t)ng; c->rm.n = (int32_t)nr; c->rm.g_tid = c->r_gtid.as<int32_t>(); c->rm.g_off = c->r_goff.as<int32_t>();
            c->rm.start = c->r_start.as<int32_t>(); c->rm.pmax_end = c->r_pmax.as<int32_t>();
        }
    }
    CK(cudaStreamSynchronize(c->st));
    return LRB_OK;
}

// ------------------------------------------------------------------------------------------- gather + canonical merge
int lrb_update_gather(lrb_ctx *c, int64_t name_base)
{
    if (!c) return LRB_E_ARG;
    MultiState *m = c->multi;
    if (!m) return fail(c, LRB_E_ARG, "lrb_update_gather: no communicator (lrb_comm_init)");
    if (!c->have_update) return fail(c, LRB_E_ARG, "lrb_update_gather: update stage has not run");
    CK(cudaSetDevice(c->device));
    int rc; const int R = m->n_ranks, me = m->rank; const bool root = me == 0;
    const int64_t l0 = total_launches();
    m->have = false;
    CK(cudaEventRecord(m->ev[0], c->st));
    // ---- this shard's updated_T as a self-contained table
    int64_t ne = 0;
    if ((rc = build_update_table(c, &ne, true))) return rc;
    const int64_t nu = c->mg.n_out, nbed = c->last_up.want_summary ? c->n_bed : 0, nkg = c->last_up.want_summary ? c->n_kg_pairs : 0;
    // ---- counts of every rank
    int64_t *mh = m->meta_host.as<int64_t>(), *mine = mh + (size_t)me * META_WORDS;
    memset(mine, 0, META_WORDS * 8);
    mine[M_NU] = nu; mine[M_NE] = ne; mine[M_NBED] = nbed; mine[M_NKG] = nkg; mine[M_NAME_BASE] = name_base; mine[M_XLOCUS] = c->xlocus_seen ? 1 : 0;
    mine[M_PARTIAL] = c->summary[LRB_S_NOVEL_PARTIAL]; mine[M_WANT_SUMMARY] = c->last_up.want_summary;
    for (int k = 0; k < LRB_S_COUNT; ++k) mine[M_SUMMARY0 + k] = c->summary[k];
    int64_t *md = m->meta_dev.as<int64_t>();
    CK(cudaMemcpyAsync(md + (size_t)me * META_WORDS, mine, META_WORDS * 8, cudaMemcpyHostToDevice, c->st));
    NK(m->api->AllGather(md + (size_t)me * META_WORDS, md, META_WORDS, ncclInt64, m->comm, c->st));
    CK(cudaMemcpyAsync(mh, md, (size_t)R * META_WORDS * 8, cudaMemcpyDeviceToHost, c->st));
    CK(cudaStreamSynchronize(c->st));
    std::vector<int64_t> o_n(R + 1, 0), o_e(R + 1, 0), o_b(R + 1, 0), o_k(R + 1, 0);
    bool want_summary = true;
    for (int r = 0; r < R; ++r) {
        const int64_t *x = mh + (size_t)r * META_WORDS;
        o_n[r + 1] = o_n[r] + x[M_NU]; o_e[r + 1] = o_e[r] + x[M_NE]; o_b[r + 1] = o_b[r] + x[M_NBED]; o_k[r + 1] = o_k[r] + x[M_NKG];
        want_summary = want_summary && x[M_WANT_SUMMARY] != 0;
    }
    const int64_t N = o_n[R], NE = o_e[R], NB = o_b[R], NK_ = o_k[R];
    if (NE >= ((int64_t)1 << 32) || N >= ((int64_t)1 << 31)) return fail(c, LRB_E_ARG, "lrb_update_gather: gathered table exceeds 2^32 exons / 2^31 transcripts");
    // ---- gatherv: every rank sends its arrays, the root receives each at the shard's offset
    struct Arr { Buf *src; Buf *dst; int esz; const std::vector<int64_t> *off; int64_t extra; };
    if (root) {
        const size_t n1 = (size_t)std::max<int64_t>(N, 1), e1 = (size_t)std::max<int64_t>(NE, 1), b1 = (size_t)std::max<int64_t>(NB, 1), k1 = (size_t)std::max<int64_t>(NK_, 1);
        for (Buf *b : {&m->g_name, &m->g_piece, &m->g_ttid, &m->g_tstart, &m->g_tend, &m->g_etid, &m->g_cov, &m->g_ref}) NEED(*b, n1 * 4);
        NEED(m->g_off, (n1 + 1) * 4); NEED(m->g_trev, n1); NEED(m->g_erev, n1);
        NEED(m->g_es, e1 * 4); NEED(m->g_ee, e1 * 4); NEED(m->g_flag, e1);
        for (Buf *b : {&m->g_bd_tid, &m->g_bd_s, &m->g_bd_e, &m->g_bd_sc}) NEED(*b, b1 * 4);
        NEED(m->g_bd_ty, b1); NEED(m->g_bd_rv, b1); NEED(m->g_kg, k1 * 8);
    }
    const Arr arrs[] = {
        {&c->tb_name, &m->g_name, 4, &o_n}, {&c->tb_piece, &m->g_piece, 4, &o_n}, {&c->tb_ttid, &m->g_ttid, 4, &o_n}, {&c->tb_tstart, &m->g_tstart, 4, &o_n},
        {&c->tb_tend, &m->g_tend, 4, &o_n}, {&c->tb_trev, &m->g_trev, 1, &o_n}, {&c->tb_etid, &m->g_etid, 4, &o_n}, {&c->tb_erev, &m->g_erev, 1, &o_n},
        {&c->tb_cov, &m->g_cov, 4, &o_n}, {&c->tb_ref, &m->g_ref, 4, &o_n}, {&c->tb_off, &m->g_off, 4, &o_n},
        {&c->tb_es, &m->g_es, 4, &o_e}, {&c->tb_ee, &m->g_ee, 4, &o_e}, {&c->tb_flag, &m->g_flag, 1, &o_e},
        {&c->bd_tid, &m->g_bd_tid, 4, &o_b}, {&c->bd_s, &m->g_bd_s, 4, &o_b}, {&c->bd_e, &m->g_bd_e, 4, &o_b}, {&c->bd_sc, &m->g_bd_sc, 4, &o_b},
        {&c->bd_ty, &m->g_bd_ty, 1, &o_b}, {&c->bd_rv, &m->g_bd_rv, 1, &o_b}, {&c->kg_pairs, &m->g_kg, 8, &o_k},
    };
    NK(m->api->GroupStart());
    for (const Arr &a : arrs) {
        const std::vector<int64_t> &off = *a.off;
        if (!root) { const int64_t cnt = off[me + 1] - off[me]; if (cnt) NK(m->api->Send(a.src->p, (size_t)cnt * a.esz, ncclChar, 0, m->comm, c->st)); }
        else for (int r = 1; r < R; ++r) {
            const int64_t cnt = off[r + 1] - off[r];
            if (cnt) NK(m->api->Recv((char *)a.dst->p + (size_t)off[r] * a.esz, (size_t)cnt * a.esz, ncclChar, r, m->comm, c->st));
        }
    }
    NK(m->api->GroupEnd());
    if (root) for (const Arr &a : arrs) { const int64_t cnt = (*a.off)[1]; if (cnt) CK(cudaMemcpyAsync(a.dst->p, a.src->p, (size_t)cnt * a.esz, cudaMemcpyDeviceToDevice, c->st)); }
    CK(cudaEventRecord(m->ev[1], c->st));
    if (!root) {
        CK(cudaStreamSynchronize(c->st));
        CK(cudaEventElapsedTime(&m->ms_gather, m->ev[0], m->ev[1])); m->ms_merge = 0;
        c->launches_last = total_launches() - l0;
        return LRB_OK;
    }
    // ---- canonical merge on the root
    RebaseArgs ra{}; ra.n = R;
    for (int r = 0; r < R; ++r) { ra.row_end[r] = o_n[r + 1]; ra.exon_base[r] = o_e[r]; ra.name_base[r] = mh[(size_t)r * META_WORDS + M_NAME_BASE]; }
    rebase_kernel<<<(unsigned)(N / 256 + 1), 256, 0, c->st>>>(ra, m->g_off.as<uint32_t>(), m->g_name.as<uint32_t>(), N, NE);
    CK(cudaGetLastError());
    m->n = N; m->n_exon = NE; m->n_bed = want_summary ? NB : 0; m->daj_recomputed = false;
    memset(m->summary, 0, sizeof m->summary);
    int64_t partial = 0; bool any_xlocus = false;
    for (int r = 0; r < R; ++r) {
        const int64_t *x = mh + (size_t)r * META_WORDS;
        for (int k = 0; k < LRB_S_COUNT; ++k) m->summary[k] += (int32_t)x[M_SUMMARY0 + k];
        partial += x[M_PARTIAL]; any_xlocus = any_xlocus || x[M_XLOCUS] != 0;
    }
    const bool detect = c->last_up.split_trans != 0;
    if (N > 0 && (want_summary || detect)) {
        const size_t n1 = (size_t)N;
        for (Buf *b : {&m->v_ident, &m->v_zeros, &m->v_cnt, &m->v_fs, &m->v_le}) NEED(*b, n1 * 4);
        NEED(m->y_barcnt, n1 * 16); NEED(m->y_barseg, n1 * 16); NEED(m->y_genebar, n1 * 8); NEED(m->y_bedcnt, n1 * 4); NEED(m->y_bedoff, n1 * 4);
        NEED(m->y_counts, 64); NEED(m->y_nelem, 64); NEED(m->tiles, (size_t)(N / 256 + 1024) * 8 * 6);
        GatheredTable gt; gt.n = N; gt.n_exon = NE; gt.exon_off = m->g_off.as<uint32_t>(); gt.es = m->g_es.as<int32_t>(); gt.ee = m->g_ee.as<int32_t>();
        launch_table_view(gt, m->v_ident.as<uint32_t>(), m->v_zeros.as<uint32_t>(), m->v_cnt.as<uint32_t>(), m->v_fs.as<int32_t>(), m->v_le.as<int32_t>(), c->st);
        std::vector<int64_t> ends(o_n.begin() + 1, o_n.end());
        memcpy(mh + (size_t)MAX_RANKS * META_WORDS - MAX_RANKS, ends.data(), (size_t)R * 8);            // pinned staging behind the metas
        CK(cudaMemcpyAsync(m->shard_end_dev.p, mh + (size_t)MAX_RANKS * META_WORDS - MAX_RANKS, (size_t)R * 8, cudaMemcpyHostToDevice, c->st));
        SummaryArgs sa{};
        sa.rows.n = N; sa.rows.tid = m->g_etid.as<int32_t>(); sa.rows.is_rev = m->g_erev.as<uint8_t>(); sa.rows.ex_beg = m->g_off.as<uint32_t>(); sa.rows.ex_n = m->v_cnt.as<uint32_t>();
        sa.ex.n = NE; sa.ex.es = m->g_es.as<int32_t>(); sa.ex.ee = m->g_ee.as<int32_t>(); sa.ex.flag = m->g_flag.as<uint8_t>();
        sa.list.n = N; sa.list.row = m->v_ident.as<uint32_t>(); sa.list.lo = m->v_zeros.as<uint32_t>(); sa.list.cnt = m->v_cnt.as<uint32_t>(); sa.list.piece = m->g_piece.as<int32_t>();
        sa.upd.n = N; sa.upd.cand = m->v_ident.as<uint32_t>(); sa.upd.cov = m->g_cov.as<int32_t>(); sa.upd.tid = m->g_ttid.as<int32_t>(); sa.upd.start = m->g_tstart.as<int32_t>();
        sa.upd.end = m->g_tend.as<int32_t>(); sa.upd.fs = m->v_fs.as<int32_t>(); sa.upd.le = m->v_le.as<int32_t>();
        sa.ref = m->g_ref.as<int32_t>(); sa.anno_gene = c->anno.gene; sa.n_upd = N;
        sa.bar_cnt = m->y_barcnt.as<uint32_t>(); sa.bar_seg = m->y_barseg.as<uint32_t>(); sa.gene_bar = m->y_genebar.as<uint64_t>();
        sa.bed_cnt = m->y_bedcnt.as<uint32_t>(); sa.bed_off = m->y_bedoff.as<uint32_t>(); sa.counts = m->y_counts.as<uint32_t>();
        sa.shard_end = m->shard_end_dev.as<int64_t>(); sa.n_shards = R; sa.probe = detect ? 1 : 0;
        if (detect) {
            NEED(m->xs_pl, n1 * 4); NEED(m->xs_hx, n1 * 4); NEED(m->xs_cnt, 64);
            sa.xs_pl = m->xs_pl.as<uint32_t>(); sa.xs_hx = m->xs_hx.as<uint32_t>(); sa.xs_cnt = m->xs_cnt.as<uint32_t>(); sa.xs_cap = (uint32_t)std::min<int64_t>(N, 0x7fffffff);
        }
        // pieces anywhere: their tid-0 site / junction keys may coincide across shards (probe needs those elements in the table)
        const bool pieces = partial > 0 || detect;
        for (int pass = 0; pass < 2; ++pass) {
            // pass 0: gene set (+ the probes); pass 1 (only after a coincidence): gene, site and junction sets over the whole table
            sa.sets = (want_summary ? SUM_G : 0) | (pass == 1 ? SUM_DAJ : 0);
            SummaryArgs cnt_args = sa; cnt_args.sets = sa.sets | ((pieces && want_summary) ? SUM_DAJ : 0);
            CK(cudaMemsetAsync(m->y_nelem.p, 0, 64, c->st)); CK(cudaMemsetAsync(m->y_counts.p, 0, 64, c->st)); CK(cudaMemsetAsync(m->flags_dev.p, 0, 64, c->st));
            if (detect) CK(cudaMemsetAsync(m->xs_cnt.p, 0, 64, c->st));
            launch_summary_count(cnt_args, m->y_nelem.as<unsigned long long>(), c->st);
            CK(cudaMemcpyAsync(mh, m->y_nelem.p, 16, cudaMemcpyDeviceToHost, c->st));
            CK(cudaStreamSynchronize(c->st));
            const uint64_t n_elem = (uint64_t)mh[0];
            const uint64_t worst = 2 * n_elem + (uint64_t)NK_, capn = worst + worst / 4 + 1024;
            NEED(m->tab, capn * sizeof(HashSlot));
            CK(cudaMemsetAsync(m->tab.p, 0xFF, capn * sizeof(HashSlot), c->st));
            sa.tab.cap = capn; sa.tab.slots = m->tab.as<HashSlot>();
            if (pass == 0 && pieces && want_summary) launch_tid0_coincidence(sa, m->flags_dev.as<uint32_t>(), c->st);
            launch_summary_sets(sa, nullptr, 0, m->tiles.as<uint64_t>(), (uint32_t *)(m->y_nelem.as<uint8_t>() + 16), d_totals(c) + T_BED, c->st, c->st, nullptr, nullptr);
            if (want_summary && NK_ > 0) launch_pairs_distinct(sa.tab, m->g_kg.as<int2>(), NK_, m->y_counts.as<uint32_t>() + CNT_KG, c->st);
            CK(cudaGetLastError());
            uint32_t *hp = (uint32_t *)mh;
            CK(cudaMemcpyAsync(hp, m->y_counts.p, 48, cudaMemcpyDeviceToHost, c->st));
            CK(cudaMemcpyAsync(hp + 16, m->flags_dev.p, 4, cudaMemcpyDeviceToHost, c->st));
            CK(cudaStreamSynchronize(c->st));
            if (detect && hp[CNT_XLOCUS])
                return fail(c, LRB_E_XSHARD, "a split piece shares a junction with a transcript of another shard: the reference may merge them across shards "
                                             "(update_gtf.c:148 never stops a piece's back-scan); rerun this input unsharded");
            if (want_summary) {
                m->summary[LRB_S_UPD_GENES] = (int32_t)hp[CNT_G]; m->summary[LRB_S_KNOWN_GENES] = (int32_t)hp[CNT_KG];
                if (pass == 1) { m->summary[LRB_S_NOVEL_SITES] = (int32_t)(hp[CNT_D] + hp[CNT_A]); m->summary[LRB_S_NOVEL_JUNC] = (int32_t)hp[CNT_J]; m->daj_recomputed = true; }
            }
            if (pass == 0 && hp[16] == 0) break;
        }
    }
    (void)any_xlocus;
    CK(cudaEventRecord(m->ev[2], c->st));
    CK(cudaStreamSynchronize(c->st));
    CK(cudaEventElapsedTime(&m->ms_gather, m->ev[0], m->ev[1])); CK(cudaEventElapsedTime(&m->ms_merge, m->ev[1], m->ev[2]));
    m->have = true;
    c->launches_last = total_launches() - l0;
    return LRB_OK;
}

int lrb_gather_fetch(lrb_ctx *c, lrb_trans_table *tab, lrb_bed_list *bed, int32_t *summary)
{
    if (!c) return LRB_E_ARG;
    MultiState *m = c->multi;
    if (!m) return fail(c, LRB_E_ARG, "lrb_gather_fetch: no communicator (lrb_comm_init)");
    CK(cudaSetDevice(c->device));
    if (m->rank != 0) {                              // only the root holds the merged result
        if (tab) memset(tab, 0, sizeof *tab);
        if (bed) memset(bed, 0, sizeof *bed);
        if (summary) memset(summary, 0, sizeof(int32_t) * LRB_S_COUNT);
        return LRB_OK;
    }
    if (!m->have) return fail(c, LRB_E_ARG, "lrb_gather_fetch: lrb_update_gather has not run");
    if (summary) memcpy(summary, m->summary, sizeof m->summary);
    const size_t n = (size_t)m->n, ne = (size_t)m->n_exon, nb = (size_t)m->n_bed;
    auto dl = [&](PBuf &dst, const Buf &src, size_t bytes) -> int {
        NEEDP(dst, std::max<size_t>(bytes, 1));
        if (bytes) CK(cudaMemcpyAsync(dst.p, src.p, bytes, cudaMemcpyDeviceToHost, c->st));
        return LRB_OK;
    };
    int rc;
    if (tab) {
        const Buf *src[] = {&m->g_name, &m->g_piece, &m->g_ttid, &m->g_tstart, &m->g_tend, &m->g_trev, &m->g_etid, &m->g_erev, &m->g_cov, &m->g_ref};
        const int esz[] = {4, 4, 4, 4, 4, 1, 4, 1, 4, 4};
        for (int i = 0; i < 10; ++i) if ((rc = dl(m->p[i], *src[i], n * esz[i]))) return rc;
        if ((rc = dl(m->p[10], m->g_off, (n + 1) * 4))) return rc;
        if ((rc = dl(m->p[11], m->g_es, ne * 4))) return rc;
        if ((rc = dl(m->p[12], m->g_ee, ne * 4))) return rc;
    }
    if (bed) {
        const Buf *src[] = {&m->g_bd_tid, &m->g_bd_s, &m->g_bd_e, &m->g_bd_sc, &m->g_bd_ty, &m->g_bd_rv};
        const int esz[] = {4, 4, 4, 4, 1, 1};
        for (int i = 0; i < 6; ++i) if ((rc = dl(m->p[13 + i], *src[i], nb * esz[i]))) return rc;
    }
    CK(cudaStreamSynchronize(c->st));
    if (tab) {
        if (n == 0) m->p[10].as<uint32_t>()[0] = 0;
        tab->n = m->n; tab->name_idx = m->p[0].as<uint32_t>(); tab->piece = m->p[1].as<int32_t>(); tab->t_tid = m->p[2].as<int32_t>(); tab->t_start = m->p[3].as<int32_t>();
        tab->t_end = m->p[4].as<int32_t>(); tab->t_rev = m->p[5].as<uint8_t>(); tab->e_tid = m->p[6].as<int32_t>(); tab->e_rev = m->p[7].as<uint8_t>();
        tab->cov = m->p[8].as<int32_t>(); tab->ref_anno = m->p[9].as<int32_t>(); tab->exon_off = m->p[10].as<uint32_t>();
        tab->exon_start = m->p[11].as<int32_t>(); tab->exon_end = m->p[12].as<int32_t>();
    }
    if (bed) {
        bed->n = m->n_bed; bed->tid = m->p[13].as<int32_t>(); bed->start = m->p[14].as<int32_t>(); bed->end = m->p[15].as<int32_t>(); bed->score = m->p[16].as<int32_t>();
        bed->type = m->p[17].as<uint8_t>(); bed->is_rev = m->p[18].as<uint8_t>();
    }
    return LRB_OK;
}

int lrb_gather_timing(lrb_ctx *c, float *ms_gather, float *ms_merge)
{
    if (!c || !c->multi) return LRB_E_ARG;
    if (ms_gather) *ms_gather = c->multi->ms_gather;
    if (ms_merge) *ms_merge = c->multi->ms_merge;
    return LRB_OK;
}

}  // extern "C"
