"""GPU parity at BASELINE.json's sizes.

configs[1] (1 M Iso-Seq-like alignments vs 60 k genes / ~200 k transcripts, SJ table, rRNA table) runs through the whole path
-- filter -> bam2gtf -> update-gtf -s -l 3 -J 1 -j with summary/BED, the bench.py step -- and every output table is compared
bit for bit with the pinned CPU port (oracle/port, ~35 s on one core).  The ONT-like shape of configs[2] (indel-dense CIGARs,
~140 ops per read) is run at 0.5 M alignments (the port needs the time, not the GPU) the same way.  On top of the direct
comparison the size-independent properties of the path are checked on the same data: idempotence, shard invariance at locus
gaps (the multi-GPU decomposition), and the counter identities of summary.txt."""
import numpy as np
import pytest

import bench
from lr2rmats_b200 import api, cabi
from tests import oracle_port as op
from tests.test_gpu_parity import assert_dict_equal, filter_valid

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    c = api.Context(0)
    yield c
    c.close()


@pytest.mark.parametrize("shape,n_reads", [("iso", 1_000_000), ("ont", 500_000)])
def test_full_size_against_port(ctx, shape, n_reads):
    fp, ep = cabi.FilterParams.default(), cabi.ExonParams.default()
    up = cabi.UpdateParams.default(full_level=3, split_trans=1, min_sj_cnt=1, want_summary=1)
    anno, rr, reads = bench.make_workload(n_reads, int(60_000 * n_reads / 1_000_000), seed=3, ont=(shape == "ont"))
    sj = bench.make_sj_table(reads, ep)
    # CPU port
    of = op.filter(reads.soa(), rr, fp)
    kept = reads.take(of["keep_idx"])
    oex = op.bam2gtf(kept.soa(), ep)
    rc, ou = op.update(oex, anno.soa(), sj, up)
    assert rc == 0
    # CUDA path, the bench.py step
    ctx.set_anno(anno.soa()); ctx.set_rm(rr); ctx.set_sj(sj)
    ctx.upload(reads.soa())
    ctx.pipeline_run(fp, ep)
    ctx.update_run(up)
    gf, gu = ctx.filter_fetch(), ctx.update_fetch()
    assert_dict_equal(filter_valid(gf), filter_valid(of))
    ou["ex"]["read_idx"] = gu["ex"]["read_idx"]
    assert_dict_equal(gu, ou)
    assert np.array_equal(gu["ex"]["read_idx"], of["keep_idx"])
    n_upd = len(gu["updated"]["cand"])
    assert n_upd > 0.15 * len(of["keep_idx"]) and (gu["novel"]["piece"] >= 0).sum() > 0

    # idempotence: the same step again on the resident batch
    ctx.pipeline_run(fp, ep); ctx.update_run(up)
    assert_dict_equal(ctx.update_fetch(), gu)

    # counter identities of summary.txt (update_gtf.c:421-534)
    s = gu["summary"].astype(np.int64)
    S = {k: int(s[i]) for i, k in enumerate(("anno_genes", "anno_trans", "upd_genes", "novel_trans", "novel_full", "novel_partial", "novel_exons",
                                              "novel_sites", "novel_junc", "known_trans", "known_genes", "uniq_known", "novel_bam", "novel_rel",
                                              "uniq_rel", "novel_unrel", "uniq_unrel", "unrecog", "uniq_unrecog"))}
    assert S["novel_trans"] == n_upd == S["novel_full"] + S["novel_partial"]
    assert S["novel_bam"] == S["novel_rel"] + S["novel_unrel"]
    assert S["known_trans"] + S["novel_bam"] + S["unrecog"] == len(of["keep_idx"])
    assert S["known_trans"] == int(((gu["cls"] & 0x1) != 0).sum())
    assert S["uniq_known"] <= S["known_trans"] and S["uniq_rel"] <= S["novel_rel"] and S["uniq_unrel"] <= S["novel_unrel"] and S["uniq_unrecog"] <= S["unrecog"]
    assert S["novel_exons"] == len(gu["bed"]["start"])

    # shard invariance at locus gaps: 8 shards (the 8-GPU decomposition) concatenate to the unsharded updated_T
    off = oex["exon_off"].astype(np.int64)
    cuts = api.shard_cuts(oex["tid"], oex["exon_start"][off[:-1]], oex["exon_end"][off[1:] - 1], 8)
    up_ns = cabi.UpdateParams.default(full_level=3, split_trans=1, min_sj_cnt=1, want_summary=0)
    whole = ctx.update_gtf(kept.soa(), ep, up_ns)
    cov, fs, le = [], [], []
    for k in range(8):
        part = kept.take(np.arange(cuts[k], cuts[k + 1]))
        g = ctx.update_gtf(part.soa(), ep, up_ns)
        cov.append(g["updated"]["cov"]); fs.append(g["updated"]["first_start"]); le.append(g["updated"]["last_end"])
    assert np.array_equal(np.concatenate(cov), whole["updated"]["cov"])
    assert np.array_equal(np.concatenate(fs), whole["updated"]["first_start"])
    assert np.array_equal(np.concatenate(le), whole["updated"]["last_end"])
    assert np.array_equal(whole["updated"]["cov"], gu["updated"]["cov"])
