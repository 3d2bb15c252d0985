"""Bisects a stage that does not come back at a given size: the data set is generated once (files under /dev/shm), every variant runs
in its own process under a timeout.  usage: debug_big.py <reads> <timeout_s>"""
import sys, os, time, subprocess, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

D = "/dev/shm/lrb_dbg"
if len(sys.argv) > 1 and sys.argv[1] == "child":
    from lr2rmats_b200 import api, cabi
    split, summary, frac = int(sys.argv[2]), int(sys.argv[3]), float(sys.argv[4])
    ld = lambda k: np.load(os.path.join(D, k + ".npy"), mmap_mode="r")
    batch = {k: ld("b_" + k) for k in ("tid", "pos", "flag", "l_qseq", "nm", "xs", "qname_hash", "cigar_off", "cigar")}
    if frac < 1.0:
        n = int(len(batch["tid"]) * frac); off = batch["cigar_off"]
        batch = {k: (np.ascontiguousarray(v[:n]) if k not in ("cigar_off", "cigar") else v) for k, v in batch.items()}
        batch["cigar_off"] = np.ascontiguousarray(off[:n + 1]); batch["cigar"] = np.ascontiguousarray(batch["cigar"][:int(off[n])])
    batch = {k: np.ascontiguousarray(v) for k, v in batch.items()}
    ctx = api.Context(0)
    ctx.set_anno({k: np.ascontiguousarray(ld("a_" + k)) for k in ("tid", "start", "end", "is_rev", "gene", "exon_off", "exon_start", "exon_end")})
    ctx.set_rm({k: np.ascontiguousarray(ld("r_" + k)) for k in ("tid", "start", "end")})
    ctx.set_sj({k: np.ascontiguousarray(ld("s_" + k)) for k in ("tid", "don", "acc", "uniq_c", "multi_c")})
    ctx.upload(batch); ctx.timing(True)
    t = time.time(); ctx.pipeline_run(cabi.FilterParams.default(), cabi.ExonParams.default()); ctx.sync()
    print("  pipeline", round(time.time() - t, 2), flush=True)
    up = cabi.UpdateParams.default(full_level=3, split_trans=split, min_sj_cnt=1, want_summary=summary)
    t = time.time(); ctx.update_run(up); ctx.sync()
    print("  update", round(time.time() - t, 2), {k: round(v, 2) for k, v in ctx.timing_get()[0].items()}, flush=True)
    sys.exit(0)

import bench
reads_n, tmo = int(sys.argv[1]), int(sys.argv[2])
os.makedirs(D, exist_ok=True)
anno, rr = bench.make_tables(60_000)
reads, sj = bench.make_dataset(anno, rr, reads_n, range(24), seed=3)
for k, v in reads.soa().items(): np.save(os.path.join(D, "b_" + k), v)
for k, v in anno.soa().items(): np.save(os.path.join(D, "a_" + k), v)
for k, v in rr.items(): np.save(os.path.join(D, "r_" + k), v)
for k, v in sj.items(): np.save(os.path.join(D, "s_" + k), v)
print("data set ready", reads.n, flush=True)
del reads
for env, split, summary, frac in (({}, 1, 0, 1.0), ({}, 0, 1, 1.0), ({}, 0, 0, 1.0), ({"LRB_SIDE_STREAM": "0", "LRB_SUM_SPLIT": "0"}, 1, 1, 1.0), ({}, 1, 1, 0.9)):
    print("variant", env, "split", split, "summary", summary, "frac", frac, flush=True)
    try:
        p = subprocess.run([sys.executable, __file__, "child", str(split), str(summary), str(frac)], env=dict(os.environ, **env), timeout=tmo)
        print("  rc", p.returncode, flush=True)
    except subprocess.TimeoutExpired:
        print("  TIMEOUT", flush=True)
import shutil; shutil.rmtree(D, ignore_errors=True)
