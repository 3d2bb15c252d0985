// oracle/port_main.cpp -- TEST INFRASTRUCTURE ONLY.
//
// `oracle/_ref/lr2rmats_port`: the drop-in CLI core (lr2rmats_b200/host/cli.cpp) driven by the CPU restatement
// (oracle/port/lr2rmats_port.c) instead of the CUDA library.  It exists to pin the restatement -- and the host
// readers/emitters it shares with the product -- against the compiled reference binary on identical input files
// (tests/test_oracle_pin.py).  The product binary (lr2rmats_b200/host/main.cpp) never links this.
#include <algorithm>
#include <cstdio>
#include <vector>
#include "../lr2rmats_b200/host/lrb_host.h"
#include "port/lr2rmats_port.h"

namespace {
struct PortEngine {
    const lrb_anno *anno = nullptr, *rm = nullptr; const lrb_sj *sj = nullptr;
    lrb_anno anno_c, rm_c; lrb_sj sj_c;
    lrb_exon_result ex{};   // kept alive for update/unique
};

int set_tables(void *s, const lrb_anno *a, const lrb_anno *rm, const lrb_sj *sj)
{
    PortEngine *e = (PortEngine *)s;
    if (a) { e->anno_c = *a; e->anno = &e->anno_c; } else e->anno = nullptr;
    if (rm) { e->rm_c = *rm; e->rm = &e->rm_c; } else e->rm = nullptr;
    if (sj) { e->sj_c = *sj; e->sj = &e->sj_c; } else e->sj = nullptr;
    return 0;
}
int do_filter(void *s, const lrb_batch *b, const lrb_filter_params *p, lrb_filter_result *out) { return orc_filter(b, ((PortEngine *)s)->rm, p, out); }
int do_bam2gtf(void *, const lrb_batch *b, const lrb_exon_params *p, lrb_exon_result *out) { return orc_bam2gtf(b, nullptr, 0, p, out); }
void chains_view(const lrb_chains *c, lrb_exon_result *ex)
{
    ex->n_reads = c->n; ex->read_idx = nullptr; ex->tid = c->tid; ex->is_rev = c->is_rev; ex->exon_off = c->exon_off;
    ex->exon_start = c->exon_start; ex->exon_end = c->exon_end;
}
int do_update(void *s, const lrb_batch *b, const lrb_chains *c, const lrb_exon_params *ep, const lrb_update_params *up, lrb_update_result *out)
{
    PortEngine *e = (PortEngine *)s;
    static lrb_anno empty{};
    if (b) { int rc = orc_bam2gtf(b, nullptr, 0, ep, &e->ex); if (rc) return rc; } else chains_view(c, &e->ex);
    return orc_update(&e->ex, e->anno ? e->anno : &empty, e->sj, up, out);
}
int do_unique(void *s, const lrb_batch *b, const lrb_chains *c, const lrb_exon_params *ep, const lrb_update_params *up, lrb_unique_result *out)
{
    PortEngine *e = (PortEngine *)s;
    if (b) { int rc = orc_bam2gtf(b, nullptr, 0, ep, &e->ex); if (rc) return rc; } else chains_view(c, &e->ex);
    return orc_unique(&e->ex, up, out);
}
int do_bam2sj(void *, const lrb_batch *b, const uint8_t *u, const lrb_sj_params *p, lrb_sj *out) { return orc_bam2sj(b, u, p, out); }
int do_sort3(void *, const uint32_t *a, const uint32_t *b, const uint32_t *c, int64_t n, const uint32_t **perm)
{
    static std::vector<uint32_t> p; p.resize((size_t)n);
    for (int64_t i = 0; i < n; ++i) p[(size_t)i] = (uint32_t)i;
    std::stable_sort(p.begin(), p.end(), [&](uint32_t x, uint32_t y) { return a[x] != a[y] ? a[x] < a[y] : b[x] != b[y] ? b[x] < b[y] : c[x] < c[y]; });
    *perm = p.data(); return 0;
}
const char *err(void *) { return "oracle port error"; }
}  // namespace

int main(int argc, char **argv)
{
    PortEngine pe; lrb::Engine eng;
    eng.self = &pe; eng.set_tables = set_tables; eng.filter = do_filter; eng.bam2gtf = do_bam2gtf; eng.update = do_update;
    eng.unique = do_unique; eng.bam2sj = do_bam2sj; eng.sort3 = do_sort3; eng.error = err;
    return lrb::cli_main(argc, argv, eng);
}
