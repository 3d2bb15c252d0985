"""N>1 host logic on CPU (gloo, world_size 2): every rank takes its locus-aligned shard of the sorted read stream
(lrb_shard_cuts), tables are broadcast from rank 0, shards are processed independently and the gathered per-shard results
concatenate to the unsharded result.  The per-shard compute here is the CPU oracle -- the CUDA path is
exercised shard by shard in tests/test_gpu_parity.py::test_shard_invariance."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from lr2rmats_b200 import api, cabi, synth
    from tests import oracle_port as op
    anno = synth.make_annotation(300, n_chrom=3, seed=41)
    reads = synth.make_reads(anno, 3000, seed=42, ont=False, reject_frac=0.0, quirk_frac=0.02)   # same seed: same stream on every rank
    # tables travel by broadcast from rank 0 (here: the annotation arrays)
    soa = anno.soa()
    for k in sorted(soa):
        t = torch.from_numpy(np.ascontiguousarray(soa[k]).view(np.uint8).copy()) if rank == 0 else torch.empty(soa[k].nbytes, dtype=torch.uint8)
        dist.broadcast(t, 0)
        soa[k] = t.numpy().view(soa[k].dtype)
    ep, up = cabi.ExonParams.default(), cabi.UpdateParams.default(full_level=3, want_summary=1)
    ex = op.bam2gtf(reads.soa(), ep)
    off = ex["exon_off"].astype(np.int64)
    cuts = api.shard_cuts(ex["tid"], ex["exon_start"][off[:-1]], ex["exon_end"][off[1:] - 1], world)
    part = reads.take(np.arange(cuts[rank], cuts[rank + 1]))
    rc, res = op.update(op.bam2gtf(part.soa(), ep), soa, None, up)
    assert rc == 0
    # gather: counts first, then padded tables (the NCCL path does the same with all_gather)
    mine = torch.from_numpy(res["updated"]["cov"].astype(np.int64))
    sizes = [torch.zeros(1, dtype=torch.int64) for _ in range(world)]
    dist.all_gather(sizes, torch.tensor([mine.numel()]))
    mx = int(max(s.item() for s in sizes))
    pad = torch.zeros(mx, dtype=torch.int64); pad[: mine.numel()] = mine
    bufs = [torch.zeros(mx, dtype=torch.int64) for _ in range(world)]
    dist.all_gather(bufs, pad)
    summ = torch.from_numpy(res["summary"].astype(np.int64)); dist.all_reduce(summ)
    if rank == 0:
        cov = np.concatenate([b[: int(s.item())].numpy() for b, s in zip(bufs, sizes)])
        rc, whole = op.update(ex, soa, None, up)
        q.put((np.array_equal(cov, whole["updated"]["cov"]), summ.numpy().tolist(), whole["summary"].tolist(), cuts.tolist()))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_shards_concatenate_exactly():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    ok, summ, whole, cuts = q.get(timeout=300)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert 0 < cuts[1] < cuts[2]
    assert ok, "concatenated per-shard updated_T differs from the unsharded run"
    # additive counters (everything but the gene counts, which need a set union when a gene spans the cut)
    from lr2rmats_b200 import cabi
    for name, a, b in zip(cabi.S_NAMES, summ, whole):
        if name not in ("upd_genes", "known_genes", "anno_genes", "anno_trans"):
            assert a == b, (name, a, b)
