#!/usr/bin/env python
"""bench.py -- alignments/sec through filter + bam2gtf + update-gtf (BASELINE.json's metric) on N B200s of one node.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--reads R] [--genes G]

A step is one pass of the hot path over one batch of synthetic alignments (workload configs[1]: 1 M Iso-Seq-like
spliced alignments vs a 60 k-gene / ~200 k-transcript annotation, plus the STAR SJ table of configs[3] and the rRNA remove
table): fused filter + CIGAR walk (`lrb_pipeline_run`) followed by classification / SJ support / split / merge / summary
(`lrb_update_run`), all through the C ABI of liblr2rmats_b200.so.

  value   device-resident: the batch is already in HBM when the timed region starts; CUDA events on the library stream,
          L2 flushed (untimed) between steps, max over ranks.
  e2e     the same step through the one-call ABI with HOST buffers: pinned H2D of the batch + the stages + D2H of every
          result table, wall clock around the call (it ends with a stream synchronise), max over ranks.
  roofline  dominant kernel: algorithmic bytes per launch / its CUDA-event duration vs MEASURED_PEAKS.json hbm_gbs.
  cpu_baseline  the UNMODIFIED reference binary (oracle/_ref/lr2rmats), single thread, on a bounded sample of the workload.

--impl reference: the reference's own CPU implementation on the host cores (one process per chromosome shard, all
cores), same metric / config, bounded sample per step.

For N > 1 launch with torchrun (the driver does): one rank per GPU, reads sharded by rank (weak scaling: every rank gets
its own batch of the same shape), the annotation / SJ / rRNA tables broadcast from rank 0 over NCCL, no data-path
collective inside the step, per-shard summary counters all-gathered after the timed region.
"""
import argparse
import ctypes as C
import json
import os
import shutil
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "long-read alignments/sec filter+bam2gtf+update-gtf"
UNIT = "alignments/s"


def log(*a):
    print(*a, file=sys.stderr, flush=True)


# ----------------------------------------------------------------------------------------------- workload
def make_workload(n_reads, n_genes, seed, ont=False):
    from lr2rmats_b200 import synth
    anno = synth.make_annotation(n_genes, n_chrom=24, seed=1)
    rr = synth.make_rrna(anno, 2000 if n_genes >= 20000 else max(10, n_genes // 30), seed=2)
    reads = synth.make_reads(anno, n_reads, seed=seed, ont=ont, reject_frac=0.2, rrna=rr)
    return anno, rr, reads


def make_sj_table(reads, ep, frac=0.7, seed=5):
    """SJ.out.tab rows from the read-derived junctions (the oracle's CIGAR walk is used here only to BUILD the input)."""
    from lr2rmats_b200 import synth
    from tests import oracle_port as op
    ex = op.bam2gtf(reads.soa(), ep)
    return synth.make_sj((ex["tid"], ex["exon_off"], ex["exon_start"], ex["exon_end"]), frac, seed=seed)


def algorithmic_bytes(reads, n_exons_rows):
    """53 + 4c + 13e bytes per alignment (c CIGAR ops, e exons)."""
    return 53 * reads.n + 4 * int(reads.cigar_off[-1]) + 13 * int(n_exons_rows)


# ------------------------------------------------------------------------------------------------ clocks
class ClockSampler:
    Q = "index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, gpu_index):
        self.rows, self.p, self.gpu = [], None, gpu_index

    def start(self):
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--id={self.gpu}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100"],
                                      stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.p = None

    def _read(self):
        for line in self.p.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.p:
            self.p.terminate()
            try:
                self.p.wait(timeout=3)
            except Exception:
                self.p.kill()
        sm = [float(r[1]) for r in self.rows if len(r) >= 9 and r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in self.rows if len(r) >= 9 and r[2].replace(".", "").isdigit()]
        reasons = set()
        for r in self.rows:
            if len(r) >= 9:
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons), "samples": len(sm)}


# ----------------------------------------------------------------------------------------- reference (CPU) arm
def reference_prepare(anno, rr, reads, sj, workdir, n_procs):
    """Writes the reference's input files (SAM text with SEQ, GTFs, SJ.out.tab), one SAM per chromosome shard."""
    from lr2rmats_b200 import synth
    n_chrom = len(anno.chrom_names)
    n_procs = max(1, min(n_procs, n_chrom))
    shards = [[] for _ in range(n_procs)]
    per_chrom = np.bincount(reads.tid[reads.tid >= 0], minlength=n_chrom)
    load = [0] * n_procs
    for c in np.argsort(-per_chrom):                       # LPT packing of chromosomes onto processes
        k = int(np.argmin(load)); shards[k].append(int(c)); load[k] += int(per_chrom[c])
    synth.write_gtf(os.path.join(workdir, "anno.gtf"), anno)
    synth.write_rm_gtf(os.path.join(workdir, "rm.gtf"), rr, anno.chrom_names)
    synth.write_sj(os.path.join(workdir, "sj.tab"), sj, anno.chrom_names)
    for k, chroms in enumerate(shards):
        idx = np.nonzero(np.isin(reads.tid, chroms))[0]
        synth.write_sam(os.path.join(workdir, f"in{k}.sam"), reads.take(idx), with_seq=True)
    return n_procs


def reference_exec(workdir, n_procs):
    """filter -> update-gtf (pass 2: -s -l 3 -J 1 -j SJ, with summary and BED) with the UNMODIFIED reference binary, one
    process per chromosome shard.  Returns wall seconds."""
    from tests import oracle_port as op
    script = ("set -e; cd {w}; {ref} filter -r rm.gtf in{k}.sam > f{k}.bam 2>/dev/null; "
              "{ref} update-gtf -s -l 3 -J 1 -j sj.tab f{k}.bam anno.gtf -y s{k}.txt -E e{k}.bed -o u{k}.gtf 2>/dev/null")
    t0 = time.perf_counter()
    procs = [subprocess.Popen(["bash", "-c", script.format(w=workdir, ref=op.REF_BIN, k=k)]) for k in range(n_procs)]
    rcs = [p.wait() for p in procs]
    dt = time.perf_counter() - t0
    if any(rcs):
        raise RuntimeError(f"reference binary failed: {rcs}")
    return dt


def bench_reference(args, rank, world):
    from tests import oracle_port as op
    from lr2rmats_b200 import cabi
    if rank != 0:
        return
    if not op.have_ref_bin():
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/lr2rmats was not built in the build container"}))
        return
    n_sample = args.ref_reads
    n_genes = max(200, int(args.genes * n_sample / args.reads))
    anno, rr, reads = make_workload(n_sample, n_genes, seed=3)
    sj = make_sj_table(reads, cabi.ExonParams.default())
    cores = os.cpu_count() or 1
    times = []
    wd = tempfile.mkdtemp(prefix="lrb_ref_")
    try:
        used = reference_prepare(anno, rr, reads, sj, wd, cores)
        for it in range(args.warmup + args.steps):
            dt = reference_exec(wd, used)
            if it >= args.warmup:
                times.append(dt)
    finally:
        shutil.rmtree(wd, ignore_errors=True)
    ms = 1e3 * float(np.mean(times))
    val = reads.n / (ms / 1e3)
    out = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
           "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int32", "data": "synthetic",
           "config": {"workload": f"configs[1] shape, bounded sample: {reads.n} Iso-Seq-like alignments vs {anno.n_genes} genes / {anno.n_trans} transcripts, "
                                  f"{len(sj['tid'])} SJ rows; reference CLI filter -> update-gtf -s -l 3 -J 1 -j (summary+BED), SAM text in",
                      "sample_reads": int(reads.n)},
           "cpu_baseline": {"value": val, "unit": UNIT, "cores": used, "kind": "reference",
                            "sample": f"{reads.n} alignments, one reference process per chromosome shard on {used} of {cores} cores"},
           "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
    print(json.dumps(out))


# --------------------------------------------------------------------------------------------------- our arm
def pinned_copy(arr):
    import torch
    t = torch.from_numpy(np.ascontiguousarray(arr)).pin_memory()
    return t, t.numpy()


def bench_ours(args, rank, world, local_rank):
    import torch
    from lr2rmats_b200 import api, cabi
    dist = None
    if world > 1:
        os.environ.setdefault("NCCL_DEBUG", "WARN")         # keeps NCCL's version banner off stdout (the JSON line is the only stdout)
        import torch.distributed as dist_
        dist = dist_
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    dev = local_rank
    fp, ep = cabi.FilterParams.default(), cabi.ExonParams.default()
    up = cabi.UpdateParams.default(full_level=3, split_trans=1, min_sj_cnt=1, want_summary=1)

    # ---- workload: every rank draws its own reads (weak scaling); tables come from rank 0 over NCCL
    t0 = time.time()
    anno, rr, reads = make_workload(args.reads, args.genes, seed=3 + 1000 * rank)
    sj = make_sj_table(reads, ep)
    tables = {"anno": anno.soa(), "rm": rr, "sj": sj}
    if world > 1:
        for name in ("anno", "rm", "sj"):
            for k in sorted(tables[name]):
                a = np.ascontiguousarray(tables[name][k])
                n = torch.tensor([a.nbytes], device="cuda", dtype=torch.int64)
                dist.broadcast(n, 0)
                buf = torch.from_numpy(a.view(np.uint8).copy()).cuda() if rank == 0 else torch.empty(int(n.item()), dtype=torch.uint8, device="cuda")
                dist.broadcast(buf, 0)                        # replicated tables: NCCL broadcast over NVLink
                tables[name][k] = buf.cpu().numpy().view(a.dtype)
    log(f"[rank {rank}] workload ready in {time.time() - t0:.1f}s: {reads.n} alignments, {int(reads.cigar_off[-1])} CIGAR ops, "
        f"{len(tables['anno']['tid'])} transcripts, {len(tables['sj']['tid'])} SJ rows")

    ctx = api.Context(dev)
    ctx.set_anno(tables["anno"]); ctx.set_rm(tables["rm"]); ctx.set_sj(tables["sj"])
    soa = reads.soa()
    keep = []
    pinned = {}
    for k, v in soa.items():
        t, a = pinned_copy(v); keep.append(t); pinned[k] = a
    batch, bk = cabi.make_batch(pinned)
    h2d_bytes = int(sum(a.nbytes for a in pinned.values()))
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=f"cuda:{dev}")

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize(dev); ctx.sync()

    def step_resident():
        ctx.pipeline_run(fp, ep)
        ctx.update_run(up)

    # ---- device-resident timing
    ctx.upload_struct(batch); ctx.sync()
    ctx.timing(True)
    for _ in range(args.warmup):
        step_resident()
    clocks = ClockSampler(dev); clocks.start()
    times, stage_ms, launches = [], [], 0
    barrier()
    for _ in range(args.steps):
        flush.fill_(1); torch.cuda.synchronize(dev)           # L2 flush, untimed
        l0 = ctx.launch_count()
        ctx.mark(0)
        step_resident()
        ctx.mark(1)
        times.append(ctx.elapsed_ms(0, 1))
        launches += ctx.launch_count() - l0
        stage_ms.append(ctx.timing_get()[0])
    barrier()
    ms_step = float(np.mean(times))
    res = ctx.update_fetch()
    nr, ne = int(res["ex"]["n_reads"]), int(res["ex"]["exon_off"][-1])
    n_novel_cand = len(res["novel"]["read"])
    summary = res["summary"].copy()
    del res

    # ---- end to end through the C ABI with HOST buffers: every step uploads its batch from pinned host memory, runs the
    # stages and fetches every result table back to (library-owned, pinned) host memory.  `--e2e-contexts` lrb contexts
    # (one CUDA stream each, one host thread each) keep that many steps in flight so the PCIe copies of one step overlap
    # the kernels of another -- the way a multi-batch caller (the CLI on a large BAM) drives the library.
    def step_e2e(cx):
        cx.upload_struct(batch)
        cx.pipeline_run(fp, ep)
        cx.update_run(up)
        if args.e2e_fetch == "full":                  # every per-read table (what -A/-a/-k/-v/-u would print as well)
            f = cabi.FilterResult(); cx._ck(cx.L.lrb_filter_fetch(cx.h, C.byref(f)))
            r = cx.update_fetch(raw=True)
            nr = int(r.ex.n_reads); ne = int(r.ex.exon_off[nr]) if nr else 0
            return int(f.n * 9 + f.n_keep * 4 + nr * (4 + 4 + 1 + 4 + 4 + 4) + 4 + ne * 9 + (r.n_known + r.n_unrecog) * 4 + r.novel.n * 16 + r.updated.n * 28 + r.bed.n * 18)
        # the outputs the reference arm's command writes: filter's kept records, updated GTF rows, BED rows, summary counters
        nk, _ = cx.filter_fetch_keep(raw=True)
        t, b, _s = cx.update_fetch_table(raw=True)
        nt = int(t.n); nte = int(t.exon_off[nt]) if nt else 0
        return int(nk * 4 + nt * (4 * 9 + 2) + 4 + nte * 8 + b.n * 18 + 19 * 4)

    ctx.timing(False)
    n_ctx = max(1, args.e2e_contexts)
    ctxs = [ctx]
    for _ in range(n_ctx - 1):
        cx = api.Context(dev)
        cx.set_anno(tables["anno"]); cx.set_rm(tables["rm"]); cx.set_sj(tables["sj"])
        ctxs.append(cx)
    last = {}

    def worker(k, n_steps):
        for _ in range(n_steps):
            last[k] = step_e2e(ctxs[k])

    def run_e2e(n_steps):
        share = [n_steps // n_ctx + (1 if k < n_steps % n_ctx else 0) for k in range(n_ctx)]
        th = [threading.Thread(target=worker, args=(k, share[k])) for k in range(n_ctx) if share[k]]
        t1 = time.perf_counter()
        for x in th:
            x.start()
        for x in th:
            x.join()
        return 1e3 * (time.perf_counter() - t1)

    run_e2e(max(n_ctx, args.warmup))
    flush.fill_(1); torch.cuda.synchronize(dev)
    barrier()
    e2e_total_ms = run_e2e(args.steps)
    barrier()
    d2h_bytes = last[0]
    e2e_ms = e2e_total_ms / args.steps
    for cx in ctxs[1:]:
        cx.sync()
    clk = clocks.stop()

    # ---- max over ranks
    n_total = reads.n
    if dist is not None:
        t = torch.tensor([ms_step, e2e_ms], device="cuda", dtype=torch.float64); dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_step, e2e_ms = float(t[0]), float(t[1])
        cnt = torch.tensor([reads.n, launches, h2d_bytes, d2h_bytes], device="cuda", dtype=torch.int64); dist.all_reduce(cnt)
        n_total, launches, h2d_bytes, d2h_bytes = (int(x) for x in cnt)
        # per-shard summary counters are gathered for the final canonical merge (plain sums; gene sets unioned on the host)
        allsum = [torch.zeros(19, dtype=torch.int32, device="cuda") for _ in range(world)]
        dist.all_gather(allsum, torch.from_numpy(summary.astype(np.int32)).cuda())
        summary = torch.stack(allsum).sum(0).cpu().numpy()
    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0)); peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if peaks else "fallback 6650 GB/s"
    st = {k: float(np.mean([s[k] for s in stage_ms])) for k in stage_ms[0]}
    n_kept = nr
    # per-kernel algorithmic bytes: scan reads 31 B/record + CIGAR, writes 9 B/record filter outputs,
    # 25 B/row + 8 B/exon; classify reads 21 B/row + 8 B/exon and writes 13 B/row + 1 B/exon
    scan_bytes = reads.n * (31 + 9) + 4 * int(reads.cigar_off[-1]) + n_kept * 25 + ne * 8
    classify_bytes = n_kept * (21 + 13) + ne * 9
    kernels = {"cigar_scan_kernel": (st["k_scan"], scan_bytes), "classify_row_kernel": (st["classify"], classify_bytes),
               "merge_fold_kernel": (st["k_fold"], n_kept * 0 + n_novel_cand * 28 + ne * 8)}
    dom = max(kernels, key=lambda k: kernels[k][0])
    dms, dbytes = kernels[dom]
    achieved = dbytes / (dms * 1e-3) / 1e9 if dms > 0 else 0.0
    path_bytes = algorithmic_bytes(reads, ne)
    traffic = None
    try:
        traffic = json.load(open(os.path.join(ROOT, "profiles", "r01_traffic.json"))).get(dom)
    except Exception:
        pass

    # ---- CPU baseline: the reference binary on a bounded sample (rank 0, N=1 only)
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        try:
            from tests import oracle_port as op
            if op.have_ref_bin():
                n_s = args.cpu_reads
                a2, rr2, rd2 = make_workload(n_s, max(200, int(args.genes * n_s / args.reads)), seed=3)
                sj2 = make_sj_table(rd2, ep)
                wd = tempfile.mkdtemp(prefix="lrb_cpu_")
                try:
                    reference_prepare(a2, rr2, rd2, sj2, wd, 1)
                    dt, n2 = reference_exec(wd, 1), rd2.n
                finally:
                    shutil.rmtree(wd, ignore_errors=True)
                cpu = {"value": n2 / dt, "unit": UNIT, "cores": 1, "kind": "reference",
                       "sample": f"{n2} alignments of the same generator vs {a2.n_genes} genes: reference `filter` + `update-gtf -s -l 3 -J 1 -j` (summary+BED), {dt:.1f}s wall, 1 thread"}
        except Exception as e:  # noqa: BLE001
            cpu = {"value": None, "unit": UNIT, "cores": 1, "kind": "reference", "sample": f"failed: {e}"}

    out = {
        "metric": METRIC, "value": n_total / (ms_step * 1e-3), "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int32", "data": "synthetic",
        "config": {"workload": f"configs[1]: {reads.n} Iso-Seq-like alignments per GPU ({int(reads.cigar_off[-1]) / reads.n:.1f} CIGAR ops, "
                               f"{ne / max(n_kept, 1):.1f} exons per kept read) vs {anno.n_genes} genes / {anno.n_trans} transcripts, "
                               f"{len(tables['sj']['tid'])} SJ rows, {len(tables['rm']['tid'])} rRNA entries; filter(-v .67 -q .75 -s .98 -r) + bam2gtf + "
                               f"update-gtf -s -l 3 -J 1 -j with summary/BED",
                   "reads_per_gpu": int(reads.n), "l2": "flushed between steps (256 MiB write, untimed)", "sharding": "one batch per rank, tables broadcast (NCCL)"},
        "e2e": {"value": n_total / (e2e_ms * 1e-3), "unit": UNIT, "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": d2h_bytes, "ms_per_step": e2e_ms,
                "contexts_in_flight": n_ctx, "fetch": args.e2e_fetch, "l2": "every step re-uploads its batch from host memory; the per-step working set of the contexts in flight exceeds L2"},
        "gpu_launches": int(launches),
        "clocks": clk,
        "roofline": {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                     "peak_source": peak_src, "algorithmic_bytes_per_launch": int(dbytes), "kernel_ms": dms,
                     "path": {"algorithmic_bytes_per_step": int(path_bytes), "achieved": path_bytes / (ms_step * 1e-3) / 1e9,
                              "frac": path_bytes / (ms_step * 1e-3) / 1e9 / peak}},
        "stage_ms": st,
        "cpu_baseline": cpu,
        "summary_counters": [int(x) for x in summary],
    }
    print(json.dumps(out))
    if dist is not None:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--reads", type=int, default=1_000_000)
    ap.add_argument("--genes", type=int, default=60_000)
    ap.add_argument("--cpu-reads", type=int, default=300_000, help="sample size of the cpu_baseline leg (about 10 s of one core; the reference's summary is quadratic, so the rate falls with the sample)")
    ap.add_argument("--ref-reads", type=int, default=100_000, help="sample size per step of --impl reference")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--e2e-fetch", default="outputs", choices=["outputs", "full"],
                    help="what the e2e leg copies back: the rows of the files the reference arm writes (updated GTF, BED, summary, kept records) or every per-read table")
    ap.add_argument("--e2e-contexts", type=int, default=3, help="lrb contexts (streams + host threads) kept in flight by the e2e leg")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        bench_reference(args, rank, world)
    else:
        bench_ours(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
