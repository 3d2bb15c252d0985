"""One rank of the sharded multi-GPU run (tests/test_multi_gpu.py spawns `world` of these, one per GPU).
usage: multi_worker.py <workload.npz> <out_dir> <rank> <world> [n_shards_per_rank=1]"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lr2rmats_b200 import api, cabi, multi  # noqa: E402


def main():
    path, out, rank, world = sys.argv[1], sys.argv[2], int(sys.argv[3]), int(sys.argv[4])
    z = np.load(path)
    batch = {k[2:]: z[k] for k in z.files if k.startswith("b_")}
    idf = os.path.join(out, "comm.id")
    if rank == 0:
        open(idf + ".tmp", "wb").write(api.comm_id()); os.rename(idf + ".tmp", idf)
    t0 = time.time()
    while not os.path.exists(idf):
        if time.time() - t0 > 120: sys.exit("no communicator id")
        time.sleep(0.05)
    ctx = api.Context(rank)
    ctx.comm_init(open(idf, "rb").read(), rank, world)
    tabs = [None, None, None]
    if rank == 0:
        tabs = [{k[2:]: z[k] for k in z.files if k.startswith(p)} for p in ("a_", "r_", "s_")]
    ctx.tables_broadcast(0, *tabs)
    cuts = multi.plan_shards(batch, world)
    fp, ep = cabi.FilterParams.default(), cabi.ExonParams.default()
    up = cabi.UpdateParams.default(full_level=int(z["p_full_level"]), split_trans=int(z["p_split"]), min_sj_cnt=1, want_summary=int(z["p_summary"]))
    shard = multi.take_shard(batch, int(cuts[rank]), int(cuts[rank + 1]))
    try:
        res = multi.run_shard(ctx, shard, int(cuts[rank]), fp, ep, up)
        code = 0
    except api.LrbError as e:
        res, code = None, e.code
    if rank == 0:
        d = {"code": np.int64(code), "cuts": cuts}
        if res is not None:
            d.update({"t_" + k: v for k, v in res["table"].items()}); d.update({"bed_" + k: v for k, v in res["bed"].items()}); d["summary"] = res["summary"]
        np.savez(os.path.join(out, "merged.npz"), **d)
    ctx.comm_destroy(); ctx.close()


if __name__ == "__main__":
    main()
