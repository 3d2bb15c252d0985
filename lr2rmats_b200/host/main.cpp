// main.cpp -- the product binary `lr2rmats-b200`: the drop-in CLI (cli.cpp) bound to the CUDA library through the C ABI.
// There is no CPU engine in this binary: without a CUDA device every subcommand fails loudly.
#include <cstdio>
#include <cstdlib>
#include <future>
#include <unistd.h>
#include "lrb_host.h"

namespace {
// The CUDA context (driver initialisation, a second or more on a multi-GPU box) is created on a helper thread while the
// subcommand decodes its input files; the first engine call joins it.
struct CudaEngine {
    lrb_ctx *ctx = nullptr; std::future<int> pending;
    lrb_ctx *get()
    {
        if (pending.valid()) {
            int rc = pending.get();
            if (rc != LRB_OK) {
                fprintf(stderr, "[lr2rmats-b200] cannot create a CUDA context (code %d): this build has no CPU fallback\n", rc);
                exit(2);
            }
        }
        return ctx;
    }
};

int set_tables(void *s, const lrb_anno *a, const lrb_anno *rm, const lrb_sj *sj)
{
    lrb_ctx *c = ((CudaEngine *)s)->get(); int rc;
    if ((rc = lrb_anno_upload(c, a))) return rc;
    if ((rc = lrb_rm_upload(c, rm))) return rc;
    return lrb_sj_upload(c, sj);
}
int do_filter(void *s, const lrb_batch *b, const lrb_filter_params *p, lrb_filter_result *o) { return lrb_filter(((CudaEngine *)s)->get(), b, p, o); }
int do_bam2gtf(void *s, const lrb_batch *b, const lrb_exon_params *p, lrb_exon_result *o) { return lrb_bam2gtf(((CudaEngine *)s)->get(), b, p, o); }
int do_sort3(void *s, const uint32_t *a, const uint32_t *b, const uint32_t *c, int64_t n, const uint32_t **perm) { return lrb_sort3(((CudaEngine *)s)->get(), a, b, c, n, perm); }
int do_bam2sj(void *s, const lrb_batch *b, const uint8_t *u, const lrb_sj_params *p, lrb_sj *o) { return lrb_bam2sj(((CudaEngine *)s)->get(), b, u, p, o); }
// LRB_SORT_INPUT=1: update-gtf takes a BAM that is NOT coordinate sorted (filter's output as it is) and sorts the rows on the
// device -- the `samtools sort` hop of the pipeline (Snakefile:90) without the extra BAM round trip
bool sort_input() { static const bool on = getenv("LRB_SORT_INPUT") && atoi(getenv("LRB_SORT_INPUT")) != 0; return on; }

int do_update(void *s, const lrb_batch *b, const lrb_chains *ch, const lrb_exon_params *ep, const lrb_update_params *up, lrb_update_result *o)
{
    lrb_ctx *c = ((CudaEngine *)s)->get(); int rc;
    if (!b) { if ((rc = lrb_chains_upload(c, ch))) return rc; }
    if (!sort_input()) return lrb_update_gtf(c, b, ep, up, o);
    if (b) { if ((rc = lrb_batch_upload(c, b))) return rc; if ((rc = lrb_exon_run(c, ep, 0))) return rc; }
    if ((rc = lrb_rows_sort(c))) return rc;
    if ((rc = lrb_update_run(c, up))) return rc;
    return lrb_update_fetch(c, o);
}
int do_update_table(void *s, const lrb_batch *b, const lrb_chains *ch, const lrb_exon_params *ep, const lrb_update_params *up,
                    lrb_trans_table *tab, lrb_bed_list *bed, int32_t *summary)
{
    lrb_ctx *c = ((CudaEngine *)s)->get(); int rc;
    if (b) { if ((rc = lrb_batch_upload(c, b))) return rc; if ((rc = lrb_exon_run(c, ep, 0))) return rc; }
    else if ((rc = lrb_chains_upload(c, ch))) return rc;
    if (sort_input() && (rc = lrb_rows_sort(c))) return rc;
    if ((rc = lrb_update_run(c, up))) return rc;
    return lrb_update_fetch_table(c, tab, bed, summary);
}
int do_unique(void *s, const lrb_batch *b, const lrb_chains *ch, const lrb_exon_params *ep, const lrb_update_params *up, lrb_unique_result *o)
{
    lrb_ctx *c = ((CudaEngine *)s)->get();
    if (!b) { int rc = lrb_chains_upload(c, ch); if (rc) return rc; }
    return lrb_unique_gtf(c, b, ep, up, o);
}
const char *err(void *s) { return lrb_last_error(((CudaEngine *)s)->get()); }
}  // namespace

int main(int argc, char **argv)
{
    CudaEngine ce; lrb::Engine eng;
    if (argc >= 3) {                                  // usage-only invocations need no device
        const char *dev = getenv("LRB_DEVICE");
        const int device = dev ? atoi(dev) : 0;
        ce.pending = std::async(std::launch::async, [&ce, device] { return lrb_ctx_create(device, &ce.ctx); });
    }
    eng.self = &ce; eng.set_tables = set_tables; eng.filter = do_filter; eng.bam2gtf = do_bam2gtf; eng.bam2sj = do_bam2sj; eng.sort3 = do_sort3; eng.update = do_update; eng.update_table = getenv("LRB_FULL_FETCH") ? nullptr : do_update_table;
    eng.unique = do_unique; eng.error = err;
    lrb::IoTrace tr;
    int rc = lrb::cli_main(argc, argv, eng);
    if (ce.pending.valid()) ce.pending.wait();        // usage errors return before any engine call
    // orderly teardown: leaving the context to the driver's process cleanup (_exit) was measured to slow down the context
    // creation of the NEXT process by more than a second (LRB_QUICK_EXIT=1 keeps that variant)
    if (getenv("LRB_QUICK_EXIT")) { fflush(NULL); tr.lap("total"); _exit(rc); }
    if (ce.ctx) lrb_ctx_destroy(ce.ctx);
    tr.lap("total");
    return rc;
}
