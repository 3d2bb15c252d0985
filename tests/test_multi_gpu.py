"""The sharded multi-GPU path (lrb_comm_* / lrb_tables_broadcast / lrb_update_gather, lr2rmats_b200/multi.py) against the
single-GPU result of the same read stream: every column of the merged updated_T table, the BED rows and all 19 summary
counters must be identical.  One process per GPU over NCCL; the N > 1 cases need N visible GPUs (`gpurun --gpus N`)."""
import os
import subprocess
import sys

import numpy as np
import pytest

from lr2rmats_b200 import api, cabi, multi, synth
from tests.test_gpu_parity import assert_dict_equal

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


def n_gpus():
    import torch
    return torch.cuda.device_count()


def workload(n_reads=120_000, n_genes=6000, seed=31, ont=False):
    anno = synth.make_annotation(n_genes, n_chrom=6, seed=seed)
    rr = synth.make_rrna(anno, 60, seed=seed + 1)
    reads = synth.make_reads(anno, n_reads, seed=seed + 2, ont=ont, reject_frac=0.2, rrna=rr)
    sj = synth.make_sj_from_reads(reads, frac=0.7, seed=seed + 3)
    return anno, rr, reads, sj


def single_gpu(anno, rr, reads, sj, up, want_bed=True):
    ctx = api.Context(0)
    ctx.set_anno(anno.soa()); ctx.set_rm(rr); ctx.set_sj(sj)
    ctx.upload(reads.soa())
    ctx.pipeline_run(cabi.FilterParams.default(), cabi.ExonParams.default())
    ctx.update_run(up)
    r = ctx.update_fetch_table(want_bed=want_bed)
    ctx.close()
    return r


@pytest.mark.parametrize("split,summary", [(1, 1), (0, 1), (1, 0)])
def test_single_rank_gather_equals_fetch(split, summary):
    """A communicator of one rank: the gather root's merge (table view, gene recount, known-gene union) over one shard."""
    anno, rr, reads, sj = workload(40_000, 2000)
    up = cabi.UpdateParams.default(full_level=3, split_trans=split, min_sj_cnt=1, want_summary=summary)
    want = single_gpu(anno, rr, reads, sj, up, want_bed=bool(summary))
    ctx = api.Context(0)
    ctx.comm_init(api.comm_id(), 0, 1)
    ctx.tables_broadcast(0, anno.soa(), rr, sj)
    got = multi.run_shard(ctx, reads.soa(), 0, cabi.FilterParams.default(), cabi.ExonParams.default(), up)
    ctx.comm_destroy(); ctx.close()
    if not summary:
        got["bed"] = None
    assert_dict_equal(got, want)
    assert len(want["table"]["cov"]) > 1000 and (not split or (want["table"]["piece"] >= 0).sum() > 0)


def run_world(tmp_path, world, anno, rr, reads, sj, split=1, summary=1, level=3):
    wl = tmp_path / "wl.npz"
    d = {"b_" + k: v for k, v in reads.soa().items()}
    d.update({"a_" + k: v for k, v in anno.soa().items()}); d.update({"r_" + k: v for k, v in rr.items()}); d.update({"s_" + k: v for k, v in sj.items()})
    d.update(p_full_level=level, p_split=split, p_summary=summary)
    np.savez(wl, **d)
    procs = [subprocess.Popen([sys.executable, os.path.join(HERE, "multi_worker.py"), str(wl), str(tmp_path), str(r), str(world)],
                              stderr=subprocess.PIPE, stdout=subprocess.PIPE) for r in range(world)]
    for p in procs:
        try:
            out, err = p.communicate(timeout=600)
        except subprocess.TimeoutExpired:
            for q in procs: q.kill()
            raise
        assert p.returncode == 0, err.decode()[-3000:]
    return np.load(tmp_path / "merged.npz")


@pytest.mark.parametrize("world", [2, 4, 8])
@pytest.mark.parametrize("shape", ["iso", "ont"])
def test_sharded_equals_single_gpu(tmp_path, world, shape):
    if n_gpus() < world:
        pytest.skip(f"needs {world} GPUs")
    anno, rr, reads, sj = workload(60_000 if shape == "ont" else 200_000, 8000, ont=(shape == "ont"))
    up = cabi.UpdateParams.default(full_level=3, split_trans=1, min_sj_cnt=1, want_summary=1)
    want = single_gpu(anno, rr, reads, sj, up)
    z = run_world(tmp_path, world, anno, rr, reads, sj)
    assert int(z["code"]) == 0
    cuts = z["cuts"]
    assert len(set(cuts.tolist())) == world + 1, "degenerate shards"
    got = dict(table={k[2:]: z[k] for k in z.files if k.startswith("t_")}, bed={k[4:]: z[k] for k in z.files if k.startswith("bed_")}, summary=z["summary"])
    assert_dict_equal(got, want)
    assert (want["table"]["piece"] >= 0).sum() > 0 and want["summary"][cabi.S_UPD_GENES if hasattr(cabi, "S_UPD_GENES") else 2] > 0


def test_cross_shard_piece_is_reported(tmp_path):
    """Split pieces whose chains also exist on another chromosome in ANOTHER shard (every read twice, on chromosomes 1-2 and on
    their clones 3-4): the reference merges such pieces across chromosomes (update_gtf.c:148), a sharded run cannot, and the gather
    says so (LRB_E_XSHARD) instead of returning a diverging table."""
    if n_gpus() < 2:
        pytest.skip("needs 2 GPUs")
    anno = synth.make_annotation(1500, n_chrom=2, seed=41)
    rr = synth.make_rrna(anno, 10, seed=42)
    reads = synth.make_reads(anno, 20_000, seed=43, reject_frac=0.2, rrna=rr)
    sj = synth.make_sj_from_reads(reads, frac=0.7, seed=44)
    a2, both, s2 = synth.clone_chromosome(anno.soa(), reads, sj, 2)

    class A:                                        # run_world only needs .soa()
        def soa(self): return a2
    z = run_world(tmp_path, 2, A(), rr, both, s2)
    assert int(z["code"]) == -8
