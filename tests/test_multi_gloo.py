"""N>1 host logic on CPU (gloo, world_size 2): the sharded driver's planning (lr2rmats_b200/multi.py: locus-aligned, qname-run-safe
cuts balanced by CIGAR ops; shard slicing; name offsets) and the canonical merge rule of the gather root, with the CPU oracle doing
the per-shard compute: tables concatenate in shard order, counters add up, and the two gene counters come from a union -- here a
plain-Python replay of add_simp_gene (update_gtf.c:175-189) over the gathered (tid, gene) columns, in the product the CUDA set kernels
over the gathered table (tests/test_multi_gpu.py checks those against one GPU).  ALL 19 summary counters must equal the unsharded run."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SHARDS_PER_RANK = 4


def replay_gene_count(t_tid, gene):
    """add_simp_gene (update_gtf.c:181-189): scan the set from its end; equal gene_id -> duplicate; first entry with a smaller tid -> insert."""
    G = []
    for t, g in zip(t_tid.tolist(), gene.tolist()):
        dup = False
        for (gt, gg) in reversed(G):
            if gg == g: dup = True; break
            if t > gt: break
        if not dup: G.append((t, g))
    return len(G)


def _gather_var(x, world):
    """all_gather of int64 vectors of different lengths (count first, then padded payload -- what the NCCL path does too)."""
    x = torch.from_numpy(np.ascontiguousarray(x).astype(np.int64))
    sizes = [torch.zeros(1, dtype=torch.int64) for _ in range(world)]
    dist.all_gather(sizes, torch.tensor([x.numel()]))
    mx = max(1, int(max(s.item() for s in sizes)))
    pad = torch.zeros(mx, dtype=torch.int64); pad[: x.numel()] = x
    bufs = [torch.zeros(mx, dtype=torch.int64) for _ in range(world)]
    dist.all_gather(bufs, pad)
    return [b[: int(s.item())].numpy() for b, s in zip(bufs, sizes)]


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from lr2rmats_b200 import cabi, multi, synth
    from tests import oracle_port as op
    anno = synth.make_annotation(300, n_chrom=3, seed=41)
    rr = synth.make_rrna(anno, 6, seed=40)
    reads = synth.make_reads(anno, 6000, seed=42, ont=False, reject_frac=0.2, rrna=rr, quirk_frac=0.02)   # same seed: same stream on every rank
    sj = synth.make_sj_from_reads(reads, frac=0.7, seed=43)
    # tables travel by broadcast from rank 0 (here: the annotation arrays)
    soa = anno.soa()
    for k in sorted(soa):
        t = torch.from_numpy(np.ascontiguousarray(soa[k]).view(np.uint8).copy()) if rank == 0 else torch.empty(soa[k].nbytes, dtype=torch.uint8)
        dist.broadcast(t, 0)
        soa[k] = t.numpy().view(soa[k].dtype)
    fp, ep = cabi.FilterParams.default(), cabi.ExonParams.default()
    up = cabi.UpdateParams.default(full_level=3, split_trans=1, min_sj_cnt=1, want_summary=1)
    batch = reads.soa()
    n_sh = world * SHARDS_PER_RANK
    cuts = multi.plan_shards(batch, n_sh)

    def run(b):
        of = op.filter(b, rr, fp)
        kept = multi.take_rows(b, of["keep_idx"])
        rc, res = op.update(op.bam2gtf(kept, ep), soa, sj, up)
        assert rc == 0
        return of, res

    cov, ttid, gene, kg, name, summ = [], [], [], [], [], np.zeros(19, np.int64)
    for s in range(rank * SHARDS_PER_RANK, (rank + 1) * SHARDS_PER_RANK):
        lo, hi = int(cuts[s]), int(cuts[s + 1])
        of, res = run(multi.take_shard(batch, lo, hi))
        u = res["updated"]; row = res["novel"]["read"][u["cand"]]
        ref = res["ref_anno"][row]
        cov.append(u["cov"]); ttid.append(u["t_tid"]); gene.append(np.where(ref >= 0, soa["gene"][np.maximum(ref, 0)], -1))
        name.append(of["keep_idx"][row].astype(np.int64) + lo)                 # record index in the WHOLE stream (name_base = lo)
        known = (res["cls"] & cabi.C_KNOWN) != 0
        kref = res["ref_anno"][known]
        kg.append(res["ex"]["tid"][known].astype(np.int64) << 32 | np.where(kref >= 0, soa["gene"][np.maximum(kref, 0)], -1).astype(np.int64) & 0xFFFFFFFF)
        summ += res["summary"]
    g_cov = _gather_var(np.concatenate(cov), world); g_tid = _gather_var(np.concatenate(ttid), world); g_gene = _gather_var(np.concatenate(gene), world)
    g_kg = _gather_var(np.concatenate(kg), world); g_name = _gather_var(np.concatenate(name), world)
    st = torch.from_numpy(summ); dist.all_reduce(st)
    if rank == 0:
        of, whole = run(batch)
        merged = st.numpy().copy()
        merged[cabi.S_NAMES.index("upd_genes")] = replay_gene_count(np.concatenate(g_tid), np.concatenate(g_gene))
        merged[cabi.S_NAMES.index("known_genes")] = len(np.unique(np.concatenate(g_kg)))
        u = whole["updated"]; row = whole["novel"]["read"][u["cand"]]
        q.put(dict(cov_ok=np.array_equal(np.concatenate(g_cov), u["cov"]), name_ok=np.array_equal(np.concatenate(g_name), of["keep_idx"][row].astype(np.int64)),
                   merged=merged.tolist(), whole=whole["summary"].tolist(), plain_sum=st.numpy().tolist(), cuts=cuts.tolist(),
                   pieces=int((whole["novel"]["piece"] >= 0).sum())))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_shards_merge_exactly():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    r = q.get(timeout=300)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    cuts = r["cuts"]
    assert all(cuts[i] < cuts[i + 1] for i in range(len(cuts) - 1)), cuts
    assert r["cov_ok"], "concatenated per-shard updated_T differs from the unsharded run"
    assert r["name_ok"], "shard-local record indices + name_base do not give the indices of the whole stream"
    assert r["pieces"] > 0
    from lr2rmats_b200 import cabi
    for name, a, b in zip(cabi.S_NAMES, r["merged"], r["whole"]):
        assert a == b, (name, a, b)
