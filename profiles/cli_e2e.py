"""Drop-in CLI, files in -> files out, wall clock: the product binary (lr2rmats-b200, CUDA) next to the reference binary
(oracle/_ref/lr2rmats) on the same synthetic SAM / GTF / SJ.out.tab, outputs compared byte for byte.  Host-bound by
construction (decode + text emission); reported separately from bench.py's device numbers.  usage: cli_e2e.py [n_reads]"""
import filecmp, gzip, json, os, subprocess, sys, tempfile, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from lr2rmats_b200 import cabi, synth
from tests import oracle_port as op

n = int(sys.argv[1]) if len(sys.argv) > 1 else 300_000
ours = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "lr2rmats_b200", "host", "lr2rmats-b200")
anno, rr, reads = bench.make_workload(n, max(200, int(60_000 * n / 1_000_000)), seed=3)
sj = bench.make_sj_table(reads, cabi.ExonParams.default())
wd = tempfile.mkdtemp(prefix="lrb_cli_")
synth.write_gtf(f"{wd}/anno.gtf", anno); synth.write_rm_gtf(f"{wd}/rm.gtf", rr, anno.chrom_names); synth.write_sj(f"{wd}/sj.tab", sj, anno.chrom_names)
synth.write_sam(f"{wd}/in.sam", reads, with_seq=True)


def stages(err_bytes):
    """the [lrb io] stage lines of LRB_IO_TRACE=1 (host decode / table readers / engine / emitters), seconds"""
    out = {}
    for line in err_bytes.decode(errors="replace").splitlines():
        if line.startswith("[lrb io]"):
            k, v = line[8:].rsplit(None, 2)[0].strip(), float(line.split()[-2])
            out[k] = round(out.get(k, 0.0) + v, 4)
    return out


def run(binary, tag, threads=None):
    env = dict(os.environ)
    env["LRB_IO_TRACE"] = "1"
    if threads:
        env["LRB_THREADS"] = str(threads)
    t = {}
    t0 = time.perf_counter()
    with open(f"{wd}/{tag}.f.bam", "wb") as f:
        p = subprocess.run([binary, "filter", "-r", f"{wd}/rm.gtf", f"{wd}/in.sam"], stdout=f, stderr=subprocess.PIPE, check=True, env=env)
    t["filter_s"] = time.perf_counter() - t0
    t["filter_stages"] = stages(p.stderr)
    t0 = time.perf_counter()
    p = subprocess.run([binary, "update-gtf", "-s", "-l", "3", "-J", "1", "-j", f"{wd}/sj.tab", f"{wd}/{tag}.f.bam", f"{wd}/anno.gtf", "-y", f"{wd}/{tag}.sum.txt",
                        "-E", f"{wd}/{tag}.bed", "-o", f"{wd}/{tag}.upd.gtf"], stderr=subprocess.PIPE, check=True, env=env)
    t["update_s"] = time.perf_counter() - t0
    t["update_stages"] = stages(p.stderr)
    t["aln_per_s"] = reads.n / (t["filter_s"] + t["update_s"])
    return t


res = {"reads": int(reads.n), "sam_bytes": os.path.getsize(f"{wd}/in.sam"), "host_cores": os.cpu_count()}
res["reference"] = run(op.REF_BIN, "ref")
run(ours, "warm")                                  # CUDA context creation / page-in, not timed
res["ours"] = run(ours, "ours")
res["ours_1thread"] = run(ours, "ours1", threads=1)
same = gzip.open(f"{wd}/ref.f.bam").read() == gzip.open(f"{wd}/ours.f.bam").read()
for ext in ("sum.txt", "bed", "upd.gtf"):
    same = same and filecmp.cmp(f"{wd}/ref.{ext}", f"{wd}/ours.{ext}", shallow=False) and filecmp.cmp(f"{wd}/ref.{ext}", f"{wd}/ours1.{ext}", shallow=False)
res["outputs_identical"] = bool(same)
res["speedup"] = res["ours"]["aln_per_s"] / res["reference"]["aln_per_s"]
print(json.dumps(res))
