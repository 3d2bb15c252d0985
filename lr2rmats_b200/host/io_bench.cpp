// io_bench.cpp -- times the host decode / encode around the hot path, no GPU involved:
//   lrb-io-bench <in.sam|in.bam> [out.bam]     decode into the SoA batch (and, with out.bam, re-encode every record)
// Prints one JSON line: records, bytes in, seconds and records/s per leg, with the thread count in use (LRB_THREADS).
#include <chrono>
#include <cstdio>
#include <numeric>
#include "lrb_host.h"

int main(int argc, char **argv)
{
    if (argc < 2) { fprintf(stderr, "usage: lrb-io-bench <in.sam|in.bam> [out.bam]\n"); return 1; }
    using clk = std::chrono::steady_clock;
    lrb::Header h; lrb::Records r; std::string err;
    r.keep_raw = argc > 2;
    auto t0 = clk::now();
    if (!lrb::read_alignments(argv[1], h, r, err)) { fprintf(stderr, "%s\n", err.c_str()); return 1; }
    const double dec = std::chrono::duration<double>(clk::now() - t0).count();
    double enc = 0;
    if (argc > 2) {
        std::vector<uint32_t> idx(r.n()); std::iota(idx.begin(), idx.end(), 0u);
        FILE *f = fopen(argv[2], "wb");
        if (!f) { fprintf(stderr, "cannot write %s\n", argv[2]); return 1; }
        t0 = clk::now();
        if (!lrb::write_bam(f, h, r, idx.data(), (int64_t)idx.size(), err)) { fprintf(stderr, "%s\n", err.c_str()); return 1; }
        fclose(f);
        enc = std::chrono::duration<double>(clk::now() - t0).count();
    }
    uint64_t x = 0;
    for (size_t i = 0; i < r.n(); ++i) x = x * 1099511628211ull + r.qhash[i] + (uint64_t)r.pos[i] + r.cigar_off[i + 1];
    printf("{\"threads\": %d, \"records\": %zu, \"cigar_ops\": %zu, \"decode_s\": %.4f, \"decode_rec_per_s\": %.0f, \"encode_s\": %.4f, \"encode_rec_per_s\": %.0f, \"digest\": \"%016llx\"}\n",
           lrb::host_threads(), r.n(), r.cigar.size(), dec, r.n() / dec, enc, enc > 0 ? r.n() / enc : 0.0, (unsigned long long)x);
    return 0;
}
