"""Profiling driver: builds the bench workload once and runs N device-resident steps (for ncu).  Not a benchmark."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from lr2rmats_b200 import api, cabi

n_reads = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
n_steps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
ont = len(sys.argv) > 3 and sys.argv[3] == "ont"
fp, ep = cabi.FilterParams.default(), cabi.ExonParams.default()
up = cabi.UpdateParams.default(full_level=3, split_trans=1, want_summary=1)
n_genes = int(sys.argv[4]) if len(sys.argv) > 4 else int(60_000 * n_reads / 1_000_000)
anno, rr, reads = bench.make_workload(n_reads, n_genes, seed=3, ont=ont)
sj = bench.make_sj_table(reads, ep)
ctx = api.Context(0)
ctx.set_anno(anno.soa()); ctx.set_rm(rr); ctx.set_sj(sj)
ctx.upload(reads.soa())
for _ in range(n_steps):
    ctx.pipeline_run(fp, ep)
    ctx.update_run(up)
ctx.sync()
print("launches", ctx.launch_count())
