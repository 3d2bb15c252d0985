"""Times lrb_rows_sort (device radix sort of the kept rows by samtools' coordinate key) on the bench workload with the qname
runs shuffled, and the update stage behind it; CUDA events on the library stream."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench
from lr2rmats_b200 import api, cabi

n_reads = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
fp, ep = cabi.FilterParams.default(), cabi.ExonParams.default()
up = cabi.UpdateParams.default(full_level=3, split_trans=1, want_summary=1)
anno, rr, reads = bench.make_workload(n_reads, int(60_000 * n_reads / 1_000_000), seed=3)
sj = bench.make_sj_table(reads, ep)
h = reads.qname_hash
starts = np.r_[0, np.nonzero(h[1:] != h[:-1])[0] + 1]; ends = np.r_[starts[1:], len(h)]
order = np.random.default_rng(1).permutation(len(starts))
lens = (ends - starts)[order]
idx = np.repeat(starts[order] - np.r_[0, np.cumsum(lens)[:-1]], lens) + np.arange(lens.sum())
sh = reads.take(idx)
ctx = api.Context(0)
ctx.set_anno(anno.soa()); ctx.set_rm(rr); ctx.set_sj(sj)
ctx.upload(sh.soa())
ts, tu = [], []
for it in range(8):
    ctx.pipeline_run(fp, ep)
    l0 = ctx.launch_count()
    ctx.mark(0); ctx.rows_sort(); ctx.mark(1)
    l1 = ctx.launch_count()
    ctx.update_run(up); ctx.mark(2); ctx.sync()
    if it >= 3:
        ts.append(ctx.elapsed_ms(0, 1)); tu.append(ctx.elapsed_ms(1, 2))
res = ctx.update_fetch()
n_rows = int(res["ex"]["n_reads"])
print(json.dumps({"rows": n_rows, "sort_ms": float(np.mean(ts)), "rows_per_s": n_rows / (np.mean(ts) * 1e-3), "launches": int(l1 - l0),
                  "bytes_per_row": "8 key + 4 index per pass (read+write), 25 permute", "update_after_sort_ms": float(np.mean(tu))}))
