"""Tuning driver: prints per-stage CUDA-event times (ms, mean of the last steps) for the current env settings."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench
from lr2rmats_b200 import api, cabi

n_reads = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
ont = len(sys.argv) > 2 and sys.argv[2] == "ont"
fp, ep = cabi.FilterParams.default(), cabi.ExonParams.default()
up = cabi.UpdateParams.default(full_level=3, split_trans=1, want_summary=1)
cache = f"/tmp/lrb_tune_{n_reads}_{int(ont)}.npz"
anno, rr, reads = bench.make_workload(n_reads, int(60_000 * n_reads / 1_000_000), seed=3, ont=ont)
sj = bench.make_sj_table(reads, ep)
ctx = api.Context(0)
ctx.set_anno(anno.soa()); ctx.set_rm(rr); ctx.set_sj(sj)
ctx.upload(reads.soa())
ctx.timing(True)
acc = []
for it in range(8):
    ctx.mark(0); ctx.pipeline_run(fp, ep); ctx.update_run(up); ctx.mark(1)
    t = ctx.timing_get()[0]; t["step"] = ctx.elapsed_ms(0, 1)
    if it >= 3: acc.append(t)
print(os.environ.get("TAG", ""), json.dumps({k: round(float(np.mean([a[k] for a in acc])), 4) for k in acc[0]}))
