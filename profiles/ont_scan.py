"""Timing of the CIGAR pass on the ONT-like shape (configs[2]): per-stage CUDA-event times, scan GB/s.  Not a benchmark line."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from lr2rmats_b200 import api, cabi, synth

n_reads = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
fp, ep = cabi.FilterParams.default(), cabi.ExonParams.default()
anno = synth.make_annotation(int(60_000 * n_reads / 1_000_000), 24, 1)
rr = synth.make_rrna(anno, 2000, seed=2)
reads = synth.make_reads(anno, n_reads, seed=3, ont=True, reject_frac=0.2, rrna=rr)
ctx = api.Context(0)
ctx.set_rm(rr)
ctx.upload(reads.soa())
ctx.timing(True)
acc = []
for it in range(10):
    ctx.pipeline_run(fp, ep)
    t = ctx.timing_get()[0]
    if it >= 3: acc.append(t)
ms = float(np.mean([a["k_scan"] for a in acc]))
nb = reads.n * 40 + 4 * int(reads.cigar_off[-1])
print(os.environ.get("TAG", ""), json.dumps({"reads": reads.n, "ops": int(reads.cigar_off[-1]), "k_scan_ms": round(ms, 4), "filter_ms": round(float(np.mean([a["filter"] for a in acc])), 4),
                                             "scan_GBs": round(nb / ms / 1e6, 1)}))
