#!/usr/bin/env python
"""Regenerates tests/golden/: small input fixtures and the outputs the UNMODIFIED reference binary produces for them.

Run in the build container (needs oracle/_ref/lr2rmats, built from /root/reference by `make -C oracle ref`):

    python tests/golden/make_golden.py

The inputs are the known-answer cases of SURVEY.md App. C (classification, "known" reachability, dedup, filter, CIGAR walk
corner cases, SJ support, annotation window break, split-piece barrier) plus two small seeded synthetic sets.  Expected
outputs are whatever the reference prints; nothing here is hand-edited.  tests/test_golden.py replays every case through
the CPU restatement (always) and through the CUDA CLI (-m gpu) and compares bytes.
"""
import gzip
import json
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
REF = os.path.join(ROOT, "oracle", "_ref", "lr2rmats")
TOY_GTF = "/root/reference/test_data/gtf/original.gtf"
TOY_RRNA = "/root/reference/test_data/gtf/rRNA.gtf"

HDR = "@HD\tVN:1.0\tSO:coordinate\n@SQ\tSN:chr1\tLN:1153837\n"


def rec(q, flag, pos, cig, nm, seq="*", chrom="chr1", extra=""):
    return f"{q}\t{flag}\t{chrom}\t{pos}\t60\t{cig}\t*\t0\t0\t{seq}\t*\tNM:i:{nm}{extra}\n"


def qlen(cig):
    n, num = 0, ""
    for ch in cig:
        if ch.isdigit():
            num += ch
        else:
            if ch in "MIS=X":
                n += int(num)
            num = ""
    return n


CASES = {}


def case(name, files, cmds):
    CASES[name] = (files, cmds)


# ---- C.1 classification / C.5 SJ support
c1 = HDR + "".join([rec("r4", 0, 500000, "100M1000N100M", 0), rec("r1", 0, 1139000, "341M438N88M883N123M892N100M", 3),
                    rec("r5", 0, 1139000, "300M", 0), rec("r2", 16, 1139100, "5S241M1409N123M892N50M", 5),
                    rec("r3", 0, 1139200, "141M2I438N88M883N100M10D23M892N60M", 10)])
ALLOUT = "-A detail.txt -y summary.txt -E novel_exon.bed -a bam.gtf -k known.gtf -v novel.gtf -u unrecog.gtf -o updated.gtf"
SMALLOUT = "-A detail.txt -y summary.txt -E novel_exon.bed -k known.gtf -o updated.gtf"
case("c1_classify", {"in.sam": c1, "anno.gtf": "@TOY", "sj1.tab": "chr1\t1139341\t1140749\t2\t2\t0\t3\t0\t40\n",
                     "sj2.tab": "chr1\t1139341\t1140749\t2\t2\t0\t0\t7\t40\nchr1\t1139867\t1140749\t2\t2\t0\t9\t0\t40\n",
                     "sj3.tab": "chr1\t100\t200\t1\t1\t0\t9\t0\t40\n", "sj4.tab": "chr1\t1141900\t1142000\t1\t1\t0\t9\t0\t40\n",
                     "sj5.tab": "chr1\t1139050\t1139150\t1\t1\t0\t9\t0\t40\n"},
     {"l3": f"update-gtf -l 3 in.sam anno.gtf {ALLOUT}", "l1": f"update-gtf -l 1 in.sam anno.gtf {ALLOUT}",
      "l5": f"update-gtf in.sam anno.gtf {ALLOUT}", "l2": f"update-gtf -l 2 in.sam anno.gtf {ALLOUT}", "l4": f"update-gtf -l 4 in.sam anno.gtf {ALLOUT}",
      "sj1": f"update-gtf -l 3 -J 1 -j sj1.tab in.sam anno.gtf {ALLOUT}", "sj1s": f"update-gtf -s -l 3 -J 1 -j sj1.tab in.sam anno.gtf {ALLOUT}",
      "sj2": f"update-gtf -l 3 -J 1 -j sj2.tab in.sam anno.gtf {ALLOUT}", "sj2m": f"update-gtf -l 3 -M x -J 1 -j sj2.tab in.sam anno.gtf {ALLOUT}",
      "sj3": f"update-gtf -l 3 -J 1 -j sj3.tab in.sam anno.gtf {ALLOUT}", "sj4": f"update-gtf -l 3 -J 1 -j sj4.tab in.sam anno.gtf {ALLOUT}",
      "sj5": f"update-gtf -s -l 3 -J 1 -j sj5.tab in.sam anno.gtf {ALLOUT}", "sjd": f"update-gtf -s -d 3 -l 3 -J 2 -j sj1.tab in.sam anno.gtf {ALLOUT}",
      "b2g": "bam2gtf in.sam > bam2gtf.gtf", "uniq": "unique-gtf in.sam > unique.gtf"})

# ---- C.1b "known" reachability
extra_line = 'chr1\tx\texon\t2000\t2500\t.\t+\t.\tgene_id "GS"; transcript_id "TS"; gene_name "GSn"; transcript_name "TSn";\n'
c1b = HDR + "".join([rec("s1", 0, 2050, "400M", 0), rec("s2", 0, 2400, "400M", 0), rec("k1", 0, 1139779, "88M883N123M892N136M", 0),
                     rec("k2", 0, 1139780, "87M883N123M892N136M", 0)])
case("c1b_known", {"in.sam": c1b, "anno.gtf": "@TOY_PREPEND:" + extra_line},
     {"l5": f"update-gtf in.sam anno.gtf {ALLOUT}", "l3": f"update-gtf -l 3 in.sam anno.gtf {ALLOUT}", "f05": f"update-gtf -f 0.2 in.sam anno.gtf {ALLOUT}"})

# ---- C.2 dedup
c2 = HDR + "".join([rec("A", 0, 1000, "100M900N100M800N100M700N100M", 0), rec("D", 0, 1000, "100M900N100M800N100M700N100M500N100M", 0),
                    rec("E", 0, 1050, "50M900N100M800N100M700N150M", 0), rec("F", 0, 1050, "50M900N100M850N50M", 0),
                    rec("B", 0, 2050, "50M800N100M700N50M", 0), rec("C", 0, 2950, "50M700N100M500N100M", 0),
                    rec("S1", 0, 9000, "300M", 0), rec("S2", 16, 9010, "300M", 0), rec("S3", 0, 9200, "300M", 0), rec("S4", 0, 20000, "100M", 0)])
case("c2_dedup", {"in.sam": c2},
     {"uniq": "unique-gtf in.sam > unique.gtf", "uniqI": "unique-gtf -I in.sam > shared.gtf", "uniqs": "unique-gtf -s in.sam > unique_s.gtf",
      "uniqD": "unique-gtf -D 20 in.sam > unique_D.gtf", "uniqd": "unique-gtf -d 60 in.sam > unique_d.gtf", "uniqf": "unique-gtf -f 0.99 in.sam > unique_f.gtf"})

# ---- C.3 filter
S100 = "A" * 100
c3h = "@HD\tVN:1.0\tSO:coordinate\n@SQ\tSN:chr1\tLN:10000000\n"
c3 = c3h + "".join([
    rec("a", 0, 1000, "100M", 0, S100), rec("b", 0, 2000, "100M", 10, S100), rec("b", 256, 2500, "100M", 11, S100),
    rec("c", 0, 3000, "100M", 10, S100), rec("c", 256, 3500, "100M", 12, S100), rec("d", 0, 4000, "40S60M", 0, S100),
    rec("e", 0, 5000, "100M", 26, S100), rec("f", 0, 6000, "100M", 25, S100), rec("g", 0, 9497000, "100M", 0, S100),
    rec("h", 0, 9497838, "100M", 0, S100), rec("i", 0, 9497839, "100M", 0, S100), rec("j", 0, 9497629, "100M", 0, S100),
    rec("k", 0, 9497630, "100M", 0, S100), rec("l", 256, 100, "100M", 5, S100), rec("l", 0, 200, "50M10D50M", 12, S100),
    f"m\t4\t*\t0\t0\t*\t*\t0\t0\t{S100}\t*\tNM:i:0\n",
    rec("n", 0, 7000, "30M100N30M200N40M", 1, S100), rec("o", 0, 8000, "10H100M", 0, S100), rec("p", 16, 8500, "20S70M10S", 2, S100),
    rec("q", 0, 8600, "50M5I45M", 6, S100, extra="\tXS:A:-\tNH:i:1"),
])
case("c3_filter", {"in.sam": c3, "rm.gtf": "@RRNA"},
     {"def": "filter -r rm.gtf in.sam > out.bam", "norm": "filter in.sam > out_norm.bam", "i1": "filter -i 1 in.sam > out_i1.bam",
      "strict": "filter -v 0.9 -q 0.9 -s 0.5 in.sam > out_strict.bam"})

# ---- C.4 CIGAR walk corner cases
c4rows = [("e1", 0, "50M51D50M"), ("e2", 0, "50M50D50M"), ("e3", 0, "50M2N50M"), ("e4", 0, "50M100N2M100N50M"), ("e5", 0, "2M100N50M"),
          ("e6", 0, "50M100N2M"), ("e7", 0, "50M100N3M100N50M"), ("e8", 0, "10S20=5X5I20M100N30M10H"), ("e9", 0, "100N50M"), ("e10", 16, "50M100N50M")]
c4 = HDR + "".join(rec(q, f, 1000, cg, 0) for q, f, cg in c4rows)
c4 += rec("e11", 16, 1000, "50M100N50M", 0, extra="\tXS:A:+") + rec("e12", 0, 1000, "50M100N50M", 0, extra="\tXS:A:-")
c4 += rec("e13", 0, 1000, "50M100N50M", 0, extra="\tXS:i:3") + rec("e14", 0, 1000, "30M2P20M100N10M60D5M70D1M", 0)
case("c4_cigar", {"in.sam": c4},
     {"b2g": "bam2gtf in.sam > bam2gtf.gtf", "b2g_e": "bam2gtf -e 1 -i 101 -t 60 in.sam > bam2gtf_e.gtf", "b2g_s": "bam2gtf -s src in.sam > bam2gtf_s.gtf"})

# ---- C.6 annotation window break
def gl(ch, s, e, g, t):
    return f'{ch}\tx\texon\t{s}\t{e}\t.\t+\t.\tgene_id "{g}"; transcript_id "{t}"; gene_name "{g}n"; transcript_name "{t}n";\n'
ta = gl("chr1", 1000, 1100, "GA", "TA") + gl("chr1", 1900, 2000, "GA", "TA")
tb = gl("chr1", 50000, 50100, "GB", "TB") + gl("chr1", 50900, 51000, "GB", "TB")
tc = gl("chr1", 1500, 1600, "GC", "TC") + gl("chr1", 2500, 2600, "GC", "TC") + gl("chr1", 2900, 3000, "GC", "TC")
tu = gl("chrUn", 10, 20, "GU", "TU")
td = gl("chr1", 60000, 60100, "GD", "TD") + gl("chr1", 60900, 61000, "GD", "TD")
c6 = HDR + rec("R", 0, 1550, "51M899N101M299N50M", 0) + rec("S", 0, 60050, "51M799N50M", 0)
case("c6_window", {"in.sam": c6, "anno1.gtf": ta + tb + tc + tu + td, "anno2.gtf": ta + tc + tu + td + tb},
     {"a1": f"update-gtf in.sam anno1.gtf {ALLOUT}", "a2": f"update-gtf in.sam anno2.gtf {ALLOUT}"})

# ---- C.7 split-piece barrier
c7 = HDR + rec("r2", 16, 1139100, "241M1409N123M892N50M", 0) + rec("r3", 0, 1139200, "141M438N88M883N100M10D23M892N60M", 0) + \
    rec("r2b", 16, 1139250, "91M1409N123M892N50M", 0)
case("c7_barrier", {"in.sam": c7, "anno.gtf": "@TOY", "sj1.tab": "chr1\t1139341\t1140749\t2\t2\t0\t3\t0\t40\n"},
     {"nosplit": f"update-gtf -l 3 -J 1 -j sj1.tab in.sam anno.gtf {ALLOUT}", "split": f"update-gtf -s -l 3 -J 1 -j sj1.tab in.sam anno.gtf {ALLOUT}"})

# ---- C.8 a split piece meets an equal chain on ANOTHER chromosome (Q14.2: a piece's back-scan never stops, update_gtf.c:148)
c8h = "@HD\tVN:1.0\tSO:coordinate\n@SQ\tSN:chr1\tLN:1153837\n@SQ\tSN:chr2\tLN:1153837\n"
c8 = c8h
for ch, sfx in (("chr1", ""), ("chr2", "x")):
    c8 += rec("r2" + sfx, 16, 1139100, "241M1409N123M892N50M", 0, chrom=ch) + rec("r3" + sfx, 0, 1139200, "141M438N88M883N100M10D23M892N60M", 0, chrom=ch) + \
        rec("r2b" + sfx, 16, 1139250, "91M1409N123M892N50M", 0, chrom=ch)
case("c8_xlocus", {"in.sam": c8, "anno.gtf": "@TOY_TWO_CHROM", "sj.tab": "chr1\t1139341\t1140749\t2\t2\t0\t3\t0\t40\nchr2\t1139341\t1140749\t2\t2\t0\t3\t0\t40\n"},
     {"split": f"update-gtf -s -l 3 -J 1 -j sj.tab in.sam anno.gtf {ALLOUT}", "nosplit": f"update-gtf -l 3 -J 1 -j sj.tab in.sam anno.gtf {ALLOUT}",
      "split_noy": "update-gtf -s -l 3 -J 1 -j sj.tab in.sam anno.gtf -o updated.gtf", "split_d": f"update-gtf -s -d 2 -l 3 -J 1 -j sj.tab in.sam anno.gtf {SMALLOUT}"})

# ---- C.9 bam2sj (parse_bam.c:896-985): proper pairs with NH tags, a genome for the motif / strand columns; single-end input gives the header only
def c9_files():
    import random
    rnd = random.Random(5)
    L = 6000
    g = [[rnd.choice("acgtACGT") for _ in range(L)] for _ in range(2)]

    def plant(c, don, acc, m):
        g[c][don - 1] = m[0]; g[c][don] = m[1]; g[c][acc - 2] = m[2]; g[c][acc - 1] = m[3]
    hdr = "@HD\tVN:1.0\tSO:coordinate\n@SQ\tSN:chrA\tLN:6000\n@SQ\tSN:chrB\tLN:6000\n"

    def r(q, flag, ch, pos, cig, extra):
        return f"{q}\t{flag}\t{ch}\t{pos}\t60\t{cig}\t=\t{pos + 200}\t300\t*\t*\tNM:i:0{extra}\n"
    rows = [r("p1", 99, "chrA", 100, "50M100N50M", "\tNH:i:1"), r("p2", 147, "chrA", 120, "30M100N20M2D10M200N40M", "\tNH:i:1"),
            r("p3", 83, "chrA", 120, "30M100N70M", "\tNH:i:3"), r("p4", 163, "chrA", 100, "50M2N48M", "\tNH:i:1"), r("s1", 0, "chrA", 100, "50M100N50M", "\tNH:i:1"),
            r("p5", 99, "chrA", 500, "10S40M300N5I30M20=10X", ""), r("p6", 99, "chrA", 400, "60M400N40M", "\tNH:i:1"), "u1\t4\t*\t0\t0\t*\t*\t0\t0\t*\t*\tNM:i:0\n",
            r("q1", 99, "chrB", 100, "50M100N50M", "\tNH:i:1"), r("q2", 147, "chrB", 90, "60M100N50M3N10M", "\tNH:i:2"), r("q3", 99, "chrB", 1000, "100M", "\tNH:i:1")]
    plant(0, 150, 249, "GTAG"); plant(1, 150, 249, "CTAC"); plant(0, 460, 859, "GCAG")
    fa = "".join(f">chr{n} test\n" + "\n".join("".join(s_[i:i + 60]) for i in range(0, L, 60)) + "\n" for n, s_ in zip("AB", g))
    se = hdr + "se1\t0\tchrA\t100\t60\t50M100N50M\t*\t0\t0\t*\t*\tNM:i:0\tNH:i:1\n"
    return {"in.sam": hdr + "".join(rows), "genome.fa": fa, "se.sam": se}
case("c9_bam2sj", c9_files(), {"sj": "bam2sj -g genome.fa in.sam > out.sj", "sj_i": "bam2sj -i 150 in.sam > out_i.sj", "se": "bam2sj se.sam > out_se.sj"})

# ---- C.10 sort_gtf.sh (the pipeline's last step, Snakefile:192): concatenated GTFs out of order, unknown chromosomes ranked by first sight,
# gene / CDS / comment lines dropped, lines in front of the first transcript, fewer and more than nine tab-separated columns
def c10_gtf():
    def t(ch, s, e, tid, extra=""):
        return f'{ch}\tlr2rmats\ttranscript\t{s}\t{e}\t.\t+\t.\tgene_id "G{tid}"; transcript_id "T{tid}";{extra}\n'

    def x(ch, s, e, tid):
        return f'{ch}\tlr2rmats\texon\t{s}\t{e}\t.\t+\t.\tgene_id "G{tid}"; transcript_id "T{tid}";\n'
    g = "# a header comment\n" + x("chr3", 5, 9, 0) + "chr1\tsrc\tgene\t1\t100000\t.\t+\t.\tgene_id \"G\";\n"
    g += t("chr2", 5000, 9000, 1) + x("chr2", 5000, 6000, 1) + x("chr2", 8000, 9000, 1)
    g += t("chr1", 7000, 9000, 2) + x("chr1", 7000, 7500, 2) + "chr1\tsrc\tCDS\t7100\t7400\t.\t+\t0\tgene_id \"G2\";\n" + x("chr1", 8000, 9000, 2)
    g += t("scaffold_9", 10, 500, 3) + x("scaffold_9", 10, 500, 3) + t("chrX", 300, 900, 4) + x("chrX", 300, 900, 4)
    g += t("chr1", 7000, 8500, 5) + x("chr1", 7000, 8500, 5) + t("chr1", 7000, 9000, 6, "\textra\tcolumns") + x("chr1", 7000, 9000, 6)
    g += "# a comment in the middle\n" + t("chrUn_1", 100, 200, 7) + "chrUn_1\tlr2rmats\texon\t100\t200\n" + t("chr10", 1, 50, 8) + x("chr10", 1, 50, 8)
    g += t("chrM", 2, 40, 9) + x("chrM", 2, 40, 9) + t("scaffold_9", 5, 20, 10) + x("scaffold_9", 5, 20, 10) + "chr2\tlr2rmats\ttranscript_like\t1\t2\t.\t+\t.\tx\n"
    g += t("chr2", 5000, 9000, 11) + x("chr2", 5000, 9000, 11) + "\n" + t("chr1", 100, 200, 12).replace("\t", " ", 2) + x("chr1", 100, 200, 12)
    return g
case("c10_sort_gtf", {"unsorted.gtf": c10_gtf()}, {"sort": "sort-gtf unsorted.gtf sorted.gtf"})


def add_synthetic():
    import numpy as np
    from lr2rmats_b200 import synth
    a = synth.make_annotation(60, n_chrom=2, seed=101)
    rr = synth.make_rrna(a, 5, seed=102)
    for name, ont, seed, n in (("syn_iso", False, 103, 220), ("syn_ont", True, 104, 60)):
        r = synth.make_reads(a, n, seed=seed, ont=ont, reject_frac=0.25, rrna=rr, quirk_frac=0.05)
        d = os.path.join(HERE, name)
        os.makedirs(d, exist_ok=True)
        synth.write_gtf(os.path.join(d, "anno.gtf"), a)
        synth.write_rm_gtf(os.path.join(d, "rm.gtf"), rr, a.chrom_names)
        synth.write_sam(os.path.join(d, "in.sam"), r, with_seq=True)
        # SJ table from the reference's own bam2gtf output
        p = subprocess.run([REF, "bam2gtf", os.path.join(d, "in.sam")], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, check=True)
        junc = set(); prev = None
        for line in p.stdout.decode().splitlines():
            f = line.split("\t")
            if f[2] == "transcript":
                prev = None; continue
            if prev is not None:
                junc.add((f[0], prev + 1, int(f[3]) - 1))
            prev = int(f[4])
        rng = np.random.default_rng(seed)
        with open(os.path.join(d, "sj.tab"), "w") as fh:
            for (ch, don, acc) in sorted(junc):
                if rng.random() < 0.7:
                    fh.write(f"{ch}\t{don}\t{acc}\t1\t1\t0\t{int(rng.integers(0, 6))}\t{int(rng.integers(0, 3))}\t30\n")
        CASES[name] = (None, {
            "filt": "filter -r rm.gtf in.sam > filt.bam", "b2g": "bam2gtf in.sam > bam2gtf.gtf", "uniq": "unique-gtf in.sam > unique.gtf",
            "p1l1": f"update-gtf -l 1 in.sam anno.gtf {SMALLOUT}", "p1l3": f"update-gtf -l 3 in.sam anno.gtf {ALLOUT}", "p1l5": f"update-gtf in.sam anno.gtf {SMALLOUT}",
            "p2": f"update-gtf -l 3 -J 1 -j sj.tab in.sam anno.gtf {SMALLOUT}", "p2s": f"update-gtf -s -l 3 -J 1 -j sj.tab in.sam anno.gtf {ALLOUT}",
            "p2sd": f"update-gtf -s -d 4 -l 3 -J 2 -j sj.tab in.sam anno.gtf {SMALLOUT}", "p2c": f"update-gtf -s -c -l 5 -J 1 -j sj.tab in.sam anno.gtf {SMALLOUT}"})


def main():
    if not os.path.exists(REF):
        sys.exit("oracle/_ref/lr2rmats is missing: run `make -C oracle ref` in the build container first")
    add_synthetic()
    toy = open(TOY_GTF).read(); rrna = open(TOY_RRNA).read()
    manifest = {}
    for name, (files, cmds) in CASES.items():
        d = os.path.join(HERE, name)
        os.makedirs(d, exist_ok=True)
        if files:
            for fn, content in files.items():
                if content == "@TOY": content = toy
                elif content == "@RRNA": content = rrna
                elif content == "@TOY_TWO_CHROM":      # the toy annotation twice: chr1 and a chr2 copy at the same coordinates with its own ids
                    import re
                    content = toy + "".join(re.sub(r'(gene_id|transcript_id) "([^"]+)"', r'\1 "\2_B"', ln).replace("chr1\t", "chr2\t", 1)
                                            for ln in toy.splitlines(True) if ln.startswith("chr1\t"))
                elif content.startswith("@TOY_PREPEND:"): content = content[len("@TOY_PREPEND:"):] + toy
                open(os.path.join(d, fn), "w").write(content)
        manifest[name] = {}
        for cname, cmd in cmds.items():
            out = os.path.join(d, "expected", cname)
            shutil.rmtree(out, ignore_errors=True); os.makedirs(out)
            stdout_to = None
            if " > " in cmd:
                cmd, stdout_to = cmd.split(" > ")
            args = [x if not x.endswith((".sam", ".gtf", ".tab", ".fa")) or x in ALLOUT.split() else os.path.join(d, x) for x in cmd.split()]
            # inputs live in the case dir, outputs go to expected/<cname>/ (cwd)
            args = []
            toks = cmd.split(); outs = set(ALLOUT.split()[1::2])
            for x in toks:
                args.append(os.path.join(d, x) if (os.path.exists(os.path.join(d, x)) and x not in outs) else x)
            with open(os.path.join(out, stdout_to) if stdout_to else os.devnull, "wb") as so:
                if args[0] == "sort-gtf":             # the reference's script, not its binary (src/sort_gtf.sh)
                    p = subprocess.run(["bash", "/root/reference/src/sort_gtf.sh"] + args[1:], cwd=out, stdout=so, stderr=subprocess.PIPE)
                else:
                    p = subprocess.run([REF] + args, cwd=out, stdout=so, stderr=subprocess.PIPE)
            if p.returncode != 0:
                sys.exit(f"{name}/{cname}: reference failed: {p.stderr.decode()[-500:]}")
            # BAM outputs: store the decompressed stream (the compressed bytes depend on zlib, the records do not)
            for fn in os.listdir(out):
                if fn.endswith(".bam"):
                    raw = gzip.open(os.path.join(out, fn)).read()
                    open(os.path.join(out, fn + ".raw"), "wb").write(raw); os.remove(os.path.join(out, fn))
            manifest[name][cname] = cmd + (f" > {stdout_to}" if stdout_to else "")
    json.dump(manifest, open(os.path.join(HERE, "manifest.json"), "w"), indent=1, sort_keys=True)
    print("golden cases:", {k: len(v) for k, v in manifest.items()})


if __name__ == "__main__":
    main()
