"""Step time against sequencing depth (same 60 k-gene annotation, more reads): per-stage CUDA-event times.  Not a benchmark line."""
import sys, os, json, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from lr2rmats_b200 import api, cabi, synth

sizes = [int(x) for x in sys.argv[1].split(",")] if len(sys.argv) > 1 else [1_000_000, 4_000_000]
ont = len(sys.argv) > 2 and sys.argv[2] == "ont"
fp, ep, up = cabi.FilterParams.default(), cabi.ExonParams.default(), cabi.UpdateParams.default()
up.split_trans = 1; up.full_level = 3; up.want_summary = 1
anno = synth.make_annotation(60_000, 24, 1)
rr = synth.make_rrna(anno, 2000, seed=2)
ctx = api.Context(0)
ctx.set_anno(anno.soa()); ctx.set_rm(rr)
for n in sizes:
    t0 = time.time()
    reads = synth.make_reads(anno, n, seed=3, ont=ont, reject_frac=0.2, rrna=rr)
    gen_s = time.time() - t0
    ctx.upload(reads.soa()); ctx.timing(True)
    acc = []
    for it in range(5):
        ctx.pipeline_run(fp, ep); t1 = ctx.timing_get()[0]
        ctx.update_run(up); t2 = ctx.timing_get()[0]
        if it >= 2: acc.append({**{k: t2[k] for k in ("classify", "merge", "summary", "k_fold")}, "filter": t1["filter"], "k_scan": t1["k_scan"]})
    m = {k: round(float(np.mean([a[k] for a in acc])), 3) for k in acc[0]}
    tot = m["filter"] + m["classify"] + m["merge"] + m["summary"]
    r = ctx.update_fetch_table(want_bed=False)
    print(json.dumps({"reads": reads.n, "ont": ont, "gen_s": round(gen_s, 1), "ms": m, "total_ms": round(tot, 3), "Maln_s": round(reads.n / tot / 1e3, 1),
                      "updated": len(r["table"]["cov"]), "summary": r["summary"].tolist()}), flush=True)
