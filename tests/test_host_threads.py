"""the multi-threaded host decode / encode (BGZF inflate by block, BAM records and SAM lines by chunk,
BGZF deflate by block group) gives the same bytes as the single-threaded code and as the reference binary (htslib) on
inputs large enough to be cut into several chunks.  The port CLI (oracle/_ref/lr2rmats_port) runs the product's own
readers and writers, so no GPU is needed here."""
import gzip
import os
import subprocess
import time

import pytest

from lr2rmats_b200 import synth
from tests import oracle_port as op

pytestmark = pytest.mark.skipif(not op.have_ref_bin(), reason="reference binary not built (oracle/_ref/lr2rmats)")


def run(binary, args, out, threads=None):
    env = dict(os.environ)
    if threads is not None:
        env["LRB_THREADS"] = str(threads)
    t0 = time.perf_counter()
    with open(out, "wb") as f:
        subprocess.run([binary] + args, stdout=f, stderr=subprocess.DEVNULL, check=True, env=env)
    return time.perf_counter() - t0


@pytest.fixture(scope="module")
def big(tmp_path_factory):
    d = tmp_path_factory.mktemp("mt")
    anno = synth.make_annotation(3000, n_chrom=6, seed=41)
    rr = synth.make_rrna(anno, 30, seed=42)
    reads = synth.make_reads(anno, 40000, seed=43, ont=False, reject_frac=0.2, rrna=rr, quirk_frac=0.02)
    synth.write_sam(d / "in.sam", reads, with_seq=True)
    synth.write_rm_gtf(d / "rm.gtf", rr, anno.chrom_names)
    assert os.path.getsize(d / "in.sam") > 8 << 20          # several SAM chunks, > 100 BGZF blocks
    return d


def test_sam_decode_and_bam_encode_threads(big):
    sam = str(big / "in.sam")
    a = ["filter", "-r", str(big / "rm.gtf"), sam]
    run(op.REF_BIN, a, big / "f_ref.bam")
    run(op.PORT_BIN, a, big / "f_t1.bam", threads=1)
    run(op.PORT_BIN, a, big / "f_t8.bam", threads=8)
    assert open(big / "f_t1.bam", "rb").read() == open(big / "f_t8.bam", "rb").read()          # same BGZF blocks, same bytes
    assert gzip.open(big / "f_ref.bam").read() == gzip.open(big / "f_t8.bam").read()
    for t in (1, 8):
        run(op.PORT_BIN, ["bam2gtf", sam], big / f"g_sam_t{t}.gtf", threads=t)
    run(op.REF_BIN, ["bam2gtf", sam], big / "g_sam_ref.gtf")
    ref = open(big / "g_sam_ref.gtf", "rb").read()
    assert ref == open(big / "g_sam_t1.gtf", "rb").read() == open(big / "g_sam_t8.gtf", "rb").read() and len(ref) > 1 << 20


def test_bam_decode_threads(big):
    if not os.path.exists(big / "f_ref.bam"):
        run(op.REF_BIN, ["filter", "-r", str(big / "rm.gtf"), str(big / "in.sam")], big / "f_ref.bam")
    outs = []
    for src in ("f_ref.bam",):                                # htslib-written BGZF in, decoded by 1 and by 8 threads
        run(op.REF_BIN, ["bam2gtf", str(big / src)], big / "g_bam_ref.gtf")
        for t in (1, 3, 8):
            run(op.PORT_BIN, ["bam2gtf", str(big / src)], big / f"g_bam_t{t}.gtf", threads=t)
            outs.append(open(big / f"g_bam_t{t}.gtf", "rb").read())
    ref = open(big / "g_bam_ref.gtf", "rb").read()
    assert all(o == ref for o in outs) and len(ref) > 1 << 20
    # filter on BAM input re-emits the raw records: byte-equal streams again
    run(op.REF_BIN, ["filter", str(big / "f_ref.bam")], big / "ff_ref.bam")
    run(op.PORT_BIN, ["filter", str(big / "f_ref.bam")], big / "ff_t8.bam", threads=8)
    assert gzip.open(big / "ff_ref.bam").read() == gzip.open(big / "ff_t8.bam").read()


def test_truncated_bgzf_and_malformed_sam(big, tmp_path):
    """A BAM cut inside a block and a SAM with a broken line deep in the file end the input there, on any thread count."""
    if not os.path.exists(big / "f_ref.bam"):
        run(op.REF_BIN, ["filter", "-r", str(big / "rm.gtf"), str(big / "in.sam")], big / "f_ref.bam")
    raw = open(big / "f_ref.bam", "rb").read()
    open(tmp_path / "cut.bam", "wb").write(raw[: len(raw) * 2 // 3])
    lines = open(big / "in.sam", "rb").read().split(b"\n")
    k = len(lines) * 3 // 4
    lines[k] = lines[k].replace(b"\t", b" ", 3)
    open(tmp_path / "bad.sam", "wb").write(b"\n".join(lines))
    for name in ("cut.bam", "bad.sam"):
        outs = []
        for t in (1, 8):
            run(op.PORT_BIN, ["bam2gtf", str(tmp_path / name)], tmp_path / f"{name}.t{t}.gtf", threads=t)
            outs.append(open(tmp_path / f"{name}.t{t}.gtf", "rb").read())
        assert outs[0] == outs[1] and len(outs[0]) > 1 << 19
