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


def test_bgzf_crc_mismatch_is_read_like_the_reference(big, tmp_path):
    """A block whose CRC field is damaged but whose deflate stream is intact: the reference's htslib (1.3) never looks at the CRC and reads
    the file to its end; the product's decoder must give the same records (it only warns), on any thread count."""
    if not os.path.exists(big / "f_ref.bam"):
        run(op.REF_BIN, ["filter", "-r", str(big / "rm.gtf"), str(big / "in.sam")], big / "f_ref.bam")
    d = bytearray(open(big / "f_ref.bam", "rb").read())
    p, blocks = 0, []
    while p + 18 <= len(d):
        bsize = (d[p + 16] | (d[p + 17] << 8)) + 1
        blocks.append((p, bsize)); p += bsize
    assert len(blocks) > 10
    for k in (3, len(blocks) // 2):
        d[blocks[k][0] + blocks[k][1] - 8] ^= 0xFF               # first byte of the block's CRC32
    open(tmp_path / "crc.bam", "wb").write(d)
    run(op.REF_BIN, ["bam2gtf", str(tmp_path / "crc.bam")], tmp_path / "ref.gtf")
    want = open(tmp_path / "ref.gtf", "rb").read()
    assert len(want) > 1 << 19
    for t in (1, 8):
        run(op.PORT_BIN, ["bam2gtf", str(tmp_path / "crc.bam")], tmp_path / f"port.t{t}.gtf", threads=t)
        assert open(tmp_path / f"port.t{t}.gtf", "rb").read() == want


def test_emitters_threads(big, tmp_path):
    """Row f-2: every text output of update-gtf / unique-gtf formatted by 8 threads == 1 thread == the reference's printers."""
    from lr2rmats_b200 import cabi
    anno = synth.make_annotation(3000, n_chrom=6, seed=41)
    synth.write_gtf(tmp_path / "anno.gtf", anno)
    files = ["detail.txt", "summary.txt", "bed", "bam.gtf", "known.gtf", "novel.gtf", "unrecog.gtf", "updated.gtf"]
    flags = ["-A", "-y", "-E", "-a", "-k", "-v", "-u", "-o"]
    outs = {}
    for who, binary, t in (("ref", op.REF_BIN, None), ("t1", op.PORT_BIN, 1), ("t8", op.PORT_BIN, 8)):
        d = tmp_path / who; d.mkdir()
        a = ["update-gtf", "-l", "5", str(big / "in.sam"), str(tmp_path / "anno.gtf")]
        for fl, fn in zip(flags, files):
            a += [fl, str(d / fn)]
        run(binary, a, d / "stdout", threads=t)
        run(binary, ["unique-gtf", str(big / "in.sam")], d / "uniq.gtf", threads=t)
        outs[who] = {fn: open(d / fn, "rb").read() for fn in files + ["uniq.gtf"]}
    for fn in files + ["uniq.gtf"]:
        assert outs["ref"][fn] == outs["t1"][fn] == outs["t8"][fn], fn
    assert len(outs["ref"]["updated.gtf"]) > 1 << 20 and len(outs["ref"]["detail.txt"]) > 1 << 20 and len(outs["ref"]["known.gtf"]) > 1 << 16


@pytest.mark.parametrize("quirk", ["plain", "comments_cds", "blank_separated", "long_line", "no_names"])
def test_gtf_reader_threads(big, tmp_path, quirk):
    """The all-threads annotation reader takes plain files only; anything else must fall back to the reference's own parsing
    (fgets 1024 + sscanf state) -- the outputs equal the reference binary's either way."""
    anno = synth.make_annotation(3000, n_chrom=6, seed=41)
    synth.write_gtf(tmp_path / "anno.gtf", anno)
    lines = open(tmp_path / "anno.gtf").read().split("\n")
    k = len(lines) // 2
    while "\texon\t" not in lines[k]:
        k += 1
    if quirk == "comments_cds":
        lines[k:k] = ["# a comment", lines[k].replace("\texon\t", "\tCDS\t"), "#another"]
    elif quirk == "blank_separated":
        lines[k] = lines[k].replace("\t", " ", 4)
    elif quirk == "long_line":
        lines[k] = lines[k] + " note \"" + "ab " * 400 + "\";"          # tokens stay below the reference's ref[100] / type[20] buffers
    elif quirk == "no_names":
        lines[k] = lines[k].split("gene_name")[0].rstrip()
    open(tmp_path / "anno.gtf", "w").write("\n".join(lines))
    outs = []
    for binary, env in ((op.REF_BIN, {}), (op.PORT_BIN, {"LRB_THREADS": "8"}), (op.PORT_BIN, {"LRB_GTF_SEQUENTIAL": "1"})):
        d = tmp_path / f"o{len(outs)}"; d.mkdir()
        e = dict(os.environ); e.update(env)
        with open(d / "updated.gtf", "wb") as f:
            p = subprocess.run([binary, "update-gtf", "-l", "5", str(big / "in.sam"), str(tmp_path / "anno.gtf"), "-y", str(d / "summary.txt"), "-k", str(d / "known.gtf")],
                               stdout=f, stderr=subprocess.DEVNULL, env=e)
        outs.append((p.returncode, open(d / "updated.gtf", "rb").read(), open(d / "summary.txt", "rb").read(), open(d / "known.gtf", "rb").read()))
    assert outs[0] == outs[1] == outs[2]
    assert outs[0][0] == 0 and len(outs[0][1]) > 1 << 20


@pytest.mark.parametrize("quirk", ["plain", "short_line", "unknown_chrom_blank"])
def test_sj_reader_threads(big, tmp_path, quirk):
    """SJ.out.tab through the all-threads reader (plain files) or the reference's sscanf loop (anything else): same
    junction support, same chromosome-id extension (gtf.c:389-449)."""
    from lr2rmats_b200 import cabi
    anno = synth.make_annotation(3000, n_chrom=6, seed=41)
    rr = synth.make_rrna(anno, 30, seed=42)
    reads = synth.make_reads(anno, 40000, seed=43, ont=False, reject_frac=0.2, rrna=rr, quirk_frac=0.02)
    ex = op.bam2gtf(reads.soa(), cabi.ExonParams.default())
    sj = synth.make_sj((ex["tid"], ex["exon_off"], ex["exon_start"], ex["exon_end"]), 0.7, seed=5)
    synth.write_gtf(tmp_path / "anno.gtf", anno); synth.write_sj(tmp_path / "sj.tab", sj, anno.chrom_names)
    lines = open(tmp_path / "sj.tab").read().split("\n")
    assert len(lines) > 20000
    k = len(lines) // 2
    if quirk == "short_line":
        lines[k] = "\t".join(lines[k].split("\t")[:5])               # later columns keep the previous line's values
    elif quirk == "unknown_chrom_blank":
        lines[k] = lines[k].replace("\t", " ", 2); lines[k + 1] = "chrUn_x\t" + lines[k + 1].split("\t", 1)[1]
    open(tmp_path / "sj.tab", "w").write("\n".join(lines))
    outs = []
    for binary, env in ((op.REF_BIN, {}), (op.PORT_BIN, {"LRB_THREADS": "8"}), (op.PORT_BIN, {"LRB_GTF_SEQUENTIAL": "1"})):
        d = tmp_path / f"o{len(outs)}"; d.mkdir()
        e = dict(os.environ); e.update(env)
        with open(d / "updated.gtf", "wb") as f:
            p = subprocess.run([binary, "update-gtf", "-s", "-l", "3", "-j", str(tmp_path / "sj.tab"), str(big / "in.sam"), str(tmp_path / "anno.gtf"),
                                "-y", str(d / "summary.txt"), "-A", str(d / "detail.txt")], stdout=f, stderr=subprocess.DEVNULL, env=e)
        outs.append((p.returncode, open(d / "updated.gtf", "rb").read(), open(d / "summary.txt", "rb").read(), open(d / "detail.txt", "rb").read()))
    assert outs[0] == outs[1] == outs[2]
    assert outs[0][0] == 0 and len(outs[0][1]) > 1 << 20


def test_io_bench_tool_digest(big):
    """lrb-io-bench (the host decode / encode timer of DESIGN 6.4) decodes to the same SoA batch on 1 and 8 threads."""
    import json
    tool = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "lr2rmats_b200", "host", "lrb-io-bench")
    if not os.path.exists(tool):
        pytest.skip("lrb-io-bench not built")
    outs = []
    for t in (1, 8):
        env = dict(os.environ, LRB_THREADS=str(t))
        p = subprocess.run([tool, str(big / "in.sam"), str(big / f"iob_t{t}.bam")], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, check=True, env=env)
        outs.append(json.loads(p.stdout))
    assert outs[0]["digest"] == outs[1]["digest"] and outs[0]["records"] == outs[1]["records"] > 40000
    assert outs[0]["threads"] == 1 and outs[1]["threads"] == 8
    assert open(big / "iob_t1.bam", "rb").read() == open(big / "iob_t8.bam", "rb").read()
