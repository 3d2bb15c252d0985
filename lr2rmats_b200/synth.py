"""Seeded synthetic workloads for lr2rmats_b200 (tests + bench): annotation, alignments, SJ.out.tab, rRNA table.

Shapes follow SURVEY.md section 8(d): genes laid out left to right per chromosome, 2-18 exons, 1-6 transcripts per
gene; reads sampled from transcripts with 5' truncation / exon skipping / shifted donors / jittered ends / soft clips;
Iso-Seq-like CIGARs (one M per exon) or ONT-like CIGARs (exons broken by short I/D), plus a share of rejects for the
filter (rRNA overlap, low identity, low coverage, secondary alignments sharing the qname).

Everything is produced as structure-of-arrays numpy buffers in the layout of include/lr2rmats_b200.h; `write_*`
helpers render the same data as SAM / GTF / SJ.out.tab text so that the reference binary can consume it.
"""
from __future__ import annotations

import numpy as np

OPS = "MIDNSHP=XB"
M, I, D, N, S, H, P, EQ, X = range(9)
MAXE = 18


def splitmix64(x: np.ndarray) -> np.ndarray:
    x = (x.astype(np.uint64) + np.uint64(0x9E3779B97F4A7C15))
    with np.errstate(over="ignore"):
        z = x
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        z = z ^ (z >> np.uint64(31))
    return z


class Annotation:
    """Transcripts in FILE order (sorted by gene position), exons ascending; mirrors lrb_anno."""

    def __init__(self):
        self.chrom_names = []
        self.tid = self.start = self.end = self.is_rev = self.gene = None
        self.exon_off = self.exon_start = self.exon_end = None
        self.gene_of_trans = None  # gene index (== interned gene id)
        self.n_genes = 0

    @property
    def n_trans(self):
        return len(self.tid)

    def soa(self):
        return dict(tid=self.tid, start=self.start, end=self.end, is_rev=self.is_rev, gene=self.gene,
                    exon_off=self.exon_off, exon_start=self.exon_start, exon_end=self.exon_end)


def make_annotation(n_genes: int, n_chrom: int = 24, seed: int = 1, single_exon_frac: float = 0.03) -> Annotation:
    rng = np.random.default_rng(seed)
    per = np.full(n_chrom, n_genes // n_chrom)
    per[: n_genes % n_chrom] += 1
    tids, starts, ends, revs, genes, eoff, es, ee = [], [], [], [], [], [0], [], []
    g = 0
    for c in range(n_chrom):
        pos = 10000
        for _ in range(per[c]):
            pos += int(rng.integers(2000, 30000))
            rev = int(rng.integers(0, 2))
            ne = 1 if rng.random() < single_exon_frac else int(rng.integers(2, MAXE + 1))
            lens = rng.integers(60, 301, ne) if ne > 1 else rng.integers(300, 1500, 1)
            gaps = rng.integers(200, 4001, max(ne - 1, 0))
            s = np.empty(ne, np.int64)
            s[0] = pos
            if ne > 1:
                s[1:] = pos + np.cumsum(lens[:-1] + gaps)
            e = s + lens - 1
            ntr = int(rng.integers(1, 7)) if ne > 2 else 1
            for t in range(ntr):
                keep = np.ones(ne, bool)
                if t > 0:
                    keep[1:-1] = rng.random(ne - 2) < 0.8
                ks, ke = s[keep], e[keep]
                tids.append(c); starts.append(int(ks[0])); ends.append(int(ke[-1])); revs.append(rev); genes.append(g)
                es.append(ks); ee.append(ke); eoff.append(eoff[-1] + len(ks))
            pos = int(e[-1])
            g += 1
    a = Annotation()
    a.chrom_names = [f"chr{i + 1}" for i in range(n_chrom)]
    a.tid = np.array(tids, np.int32); a.start = np.array(starts, np.int32); a.end = np.array(ends, np.int32)
    a.is_rev = np.array(revs, np.uint8); a.gene = np.array(genes, np.int32)
    a.exon_off = np.array(eoff, np.uint32)
    a.exon_start = np.concatenate(es).astype(np.int32); a.exon_end = np.concatenate(ee).astype(np.int32)
    a.n_genes = g
    return a


def make_rrna(anno: Annotation, n: int, seed: int = 2):
    """rRNA/remove table: n single-exon entries at random positions, sorted by (tid,start) -- the order remove_overlap
    (bam_filter.c:49-59) needs for its early exit to be harmless."""
    rng = np.random.default_rng(seed)
    pick = np.sort(rng.choice(anno.n_trans, size=min(n, anno.n_trans), replace=False))
    tid = anno.tid[pick]
    st = anno.start[pick] + rng.integers(0, 50, len(pick)).astype(np.int32)
    en = st + rng.integers(100, 400, len(pick)).astype(np.int32)
    order = np.lexsort((st, tid))
    return dict(tid=tid[order].astype(np.int32), start=st[order].astype(np.int32), end=en[order].astype(np.int32))


def _padded_exons(anno: Annotation):
    nt = anno.n_trans
    ne = np.diff(anno.exon_off.astype(np.int64))
    ps = np.zeros((nt, MAXE), np.int64); pe = np.zeros((nt, MAXE), np.int64)
    idx = np.arange(len(anno.exon_start)) - np.repeat(anno.exon_off[:-1].astype(np.int64), ne)
    row = np.repeat(np.arange(nt), ne)
    ps[row, idx] = anno.exon_start; pe[row, idx] = anno.exon_end
    return ps, pe, ne


class Reads:
    """Alignment records as SoA (lrb_batch layout) + the read-group id behind each qname."""

    def __init__(self):
        self.tid = self.pos = self.flag = self.l_qseq = self.nm = self.xs = self.qname_hash = None
        self.cigar_off = self.cigar = None
        self.qid = None  # qname == f"read{qid}"
        self.chrom_names = None
        self.chrom_lens = None

    @property
    def n(self):
        return len(self.tid)

    def soa(self):
        return dict(tid=self.tid, pos=self.pos, flag=self.flag, l_qseq=self.l_qseq, nm=self.nm, xs=self.xs,
                    qname_hash=self.qname_hash, cigar_off=self.cigar_off, cigar=self.cigar)

    def take(self, idx):
        """Sub-batch (rows idx, in that order)."""
        idx = np.asarray(idx, np.int64)
        r = Reads()
        for k in ("tid", "pos", "flag", "l_qseq", "nm", "xs", "qname_hash", "qid"):
            setattr(r, k, getattr(self, k)[idx].copy())
        lens = self.cigar_off[idx + 1].astype(np.int64) - self.cigar_off[idx].astype(np.int64)
        off = np.zeros(len(idx) + 1, np.int64); np.cumsum(lens, out=off[1:])
        src = np.repeat(self.cigar_off[idx].astype(np.int64) - off[:-1], lens) + np.arange(off[-1])
        r.cigar = self.cigar[src].copy(); r.cigar_off = off.astype(np.uint64)
        r.chrom_names, r.chrom_lens = self.chrom_names, self.chrom_lens
        return r


def make_reads(anno: Annotation, n_reads: int, seed: int = 3, ont: bool = False, reject_frac: float = 0.0,
               rrna=None, quirk_frac: float = 0.01, chrom_len: int = 150_000_000, chrom: int | None = None, qid_base: int = 0) -> Reads:
    """n_reads primary alignments (+ secondaries when reject_frac>0), coordinate sorted and name grouped.
    chrom: draw from the transcripts (and rRNA entries) of that chromosome only -- a data set can then be generated chromosome by
    chromosome (in parallel, or by the rank that owns the chromosome) and concatenated; qid_base keeps the read names distinct."""
    rng = np.random.default_rng(seed)
    ps, pe, tne = _padded_exons(anno)
    n = n_reads
    inter = rng.random(n) < 0.05
    if chrom is None:
        t = rng.integers(0, anno.n_trans, n)
    else:
        on = np.nonzero(anno.tid == chrom)[0]
        t = on[rng.integers(0, len(on), n)]
        if rrna is not None:
            m = rrna["tid"] == chrom
            rrna = {k: v[m] for k, v in rrna.items()}
    ne = tne[t]
    # exon window [a, b] of the transcript kept by the read
    trunc = (rng.random(n) < 0.40) & (ne > 2)
    a = np.where(trunc, (rng.random(n) * (ne - 1)).astype(np.int64), 0)
    b = ne - 1
    k_idx = np.arange(MAXE)[None, :]
    keep = (k_idx >= a[:, None]) & (k_idx <= b[:, None])
    skip = (rng.random(n) < 0.15) & (b - a >= 2)
    sk = a + 1 + (rng.random(n) * np.maximum(b - a - 1, 1)).astype(np.int64)
    keep &= ~(skip[:, None] & (k_idx == sk[:, None]))
    s = ps[t].copy(); e = pe[t].copy()
    # intergenic reads: 1-5 synthetic exons placed after the transcript's gene end (may overlap neighbours: fine)
    ni = int(inter.sum())
    if ni:
        nie = rng.integers(1, 6, ni)
        base = (pe[t[inter], 0] * 0 + anno.end[t[inter]] + rng.integers(300, 1500, ni))
        ls = rng.integers(80, 400, (ni, MAXE)); gp = rng.integers(150, 2500, (ni, MAXE))
        cs = base[:, None] + np.cumsum(ls + gp, axis=1) - (ls + gp)
        s[inter] = cs; e[inter] = cs + ls - 1
        keep[inter] = k_idx < nie[:, None]
        trunc = trunc & ~inter
    # shifted donor (exon end) on 15% of multi-exon reads
    cnt = keep.sum(1)
    shift = (rng.random(n) < 0.15) & (cnt >= 2) & ~inter
    first = np.argmax(keep, axis=1)
    last = MAXE - 1 - np.argmax(keep[:, ::-1], axis=1)
    dv = np.array([-9, -4, 3, 6, 12])[rng.integers(0, 5, n)]
    e[np.arange(n)[shift], first[shift]] += dv[shift]
    # jitter the two ends
    rows = np.arange(n)
    flen = e[rows, first] - s[rows, first] + 1
    exact = trunc & (rng.random(n) < 0.30)
    js = np.minimum(rng.integers(0, 41, n), np.maximum(flen - 20, 0))
    js = np.where(exact | (cnt == 1) & False, 0, js)
    s[rows, first] += js
    llen = e[rows, last] - s[rows, last] + 1
    je = np.minimum(rng.integers(0, 41, n), np.maximum(llen - 20, 0))
    e[rows, last] -= je

    tid = anno.tid[t].astype(np.int32)
    rev = anno.is_rev[t].astype(np.uint8)
    rev = np.where(inter, rng.integers(0, 2, n).astype(np.uint8), rev)

    # ---- reject classes (filter): secondaries are appended later
    cls = np.zeros(n, np.int8)  # 0 clean, 1 rRNA, 2 low identity, 3 low coverage, 4 has-secondary
    if reject_frac > 0:
        u = rng.random(n)
        cls[u < reject_frac * 0.25] = 2
        cls[(u >= reject_frac * 0.25) & (u < reject_frac * 0.5)] = 3
        cls[(u >= reject_frac * 0.5) & (u < reject_frac * 0.75)] = 4
        if rrna is not None and len(rrna["tid"]):
            rr = (u >= reject_frac * 0.75) & (u < reject_frac)
            cls[rr] = 1
            k = rng.integers(0, len(rrna["tid"]), int(rr.sum()))
            # single-exon read sitting on the rRNA locus
            tid[rr] = rrna["tid"][k]
            keep[rr] = k_idx < 1
            first[rr] = 0; last[rr] = 0
            s[rr, 0] = rrna["start"][k] + rng.integers(-30, 30, int(rr.sum()))
            e[rr, 0] = s[rr, 0] + rng.integers(150, 600, int(rr.sum()))
            inter = inter | rr

    QID_BASE[0] = int(qid_base)
    r = _assemble(rng, n, tid, rev, s, e, keep, ont, cls, quirk_frac)
    QID_BASE[0] = 0
    # ---- secondary alignments: same qname, a slightly worse / equal copy right after the primary
    if reject_frac > 0:
        sec = np.nonzero(cls == 4)[0]
        if len(sec):
            dup = r.take(sec)
            dup.flag = (dup.flag | np.uint16(256)).astype(np.uint16)
            bump = rng.integers(0, 3, len(sec))  # 0: identical score (-> both dropped), 1: tiny diff, 2: clearly worse
            dup.nm = (dup.nm + np.where(bump == 0, 0, np.where(bump == 1, 1, 40))).astype(np.int32)
            order_key = np.concatenate([np.arange(n) * 2, sec * 2 + 1])
            r = _concat(r, dup)
            order = np.argsort(order_key, kind="stable")
            r = r.take(order)
    # coordinate sort, stable so that equal-qname neighbours stay adjacent (they share tid/pos)
    order = np.lexsort((np.arange(r.n), r.pos, r.tid))
    r = r.take(order)
    r.chrom_names = list(anno.chrom_names)
    r.chrom_lens = [chrom_len] * len(anno.chrom_names)
    return r


QID_BASE = [0]          # read-name offset of the batch being assembled (make_reads' qid_base)


def _concat(a: Reads, b: Reads) -> Reads:
    r = Reads()
    for k in ("tid", "pos", "flag", "l_qseq", "nm", "xs", "qname_hash", "qid"):
        setattr(r, k, np.concatenate([getattr(a, k), getattr(b, k)]))
    r.cigar = np.concatenate([a.cigar, b.cigar])
    r.cigar_off = np.concatenate([a.cigar_off[:-1].astype(np.int64), b.cigar_off.astype(np.int64) + int(a.cigar_off[-1])]).astype(np.uint64)
    return r


def _assemble(rng, n, tid, rev, s, e, keep, ont, cls, quirk_frac) -> Reads:
    """exon table (n x MAXE, masked by keep) -> packed CIGAR pool + record fields."""
    rows, cols = np.nonzero(keep)              # row-major: exons of a read in ascending order
    cnt = keep.sum(1)
    eoff = np.zeros(n + 1, np.int64); np.cumsum(cnt, out=eoff[1:])
    xs_, xe_ = s[rows, cols], e[rows, cols]
    elen = xe_ - xs_ + 1
    elen = np.maximum(elen, 1)
    is_last = np.ones(len(rows), bool); is_last[:-1] = rows[1:] != rows[:-1]
    gap = np.zeros(len(rows), np.int64)
    gap[:-1] = xs_[1:] - (xs_[:-1] + elen[:-1])
    gap[is_last] = 0
    gap = np.where(~is_last & (gap < 20), 20, gap)   # keep introns sane if edits made exons collide

    lclip = rng.integers(0, 21, n); rclip = rng.integers(0, 21, n)
    lclip[rng.random(n) < 0.3] = 0; rclip[rng.random(n) < 0.3] = 0

    if not ont:
        # per exon: M [N]
        per_exon_ops = 1 + (~is_last).astype(np.int64)
        ooff = np.zeros(len(rows) + 1, np.int64); np.cumsum(per_exon_ops, out=ooff[1:])
        words = np.zeros(ooff[-1], np.uint32)
        words[ooff[:-1]] = (elen.astype(np.uint32) << np.uint32(4)) | np.uint32(M)
        nn = ~is_last
        words[ooff[:-1][nn] + 1] = (gap[nn].astype(np.uint32) << np.uint32(4)) | np.uint32(N)
        ops_per_read = np.add.reduceat(per_exon_ops, eoff[:-1])
        qlen_core = np.add.reduceat(elen, eoff[:-1])
        indel_bases = np.zeros(n, np.int64)
    else:
        # per exon: m runs of M separated by 1-3 bp I or D; then [N]
        m = np.maximum(1, elen // 17)
        nind = m - 1
        ind_exon = np.repeat(np.arange(len(rows)), nind)
        ind_isD = rng.random(len(ind_exon)) < 0.5
        ind_len = rng.integers(1, 4, len(ind_exon))
        dsum = np.bincount(ind_exon, weights=ind_len * ind_isD, minlength=len(rows)).astype(np.int64)
        isum = np.bincount(ind_exon, weights=ind_len * (~ind_isD), minlength=len(rows)).astype(np.int64)
        mb = elen - dsum                              # ref bases left for the M runs (>= m because elen//17 runs)
        base = mb // m; rem = mb - base * m
        per_exon_ops = 2 * m - 1 + (~is_last).astype(np.int64)
        ooff = np.zeros(len(rows) + 1, np.int64); np.cumsum(per_exon_ops, out=ooff[1:])
        words = np.zeros(ooff[-1], np.uint32)
        run_exon = np.repeat(np.arange(len(rows)), m)
        run_k = np.arange(len(run_exon)) - np.repeat(np.cumsum(m) - m, m)
        run_len = base[run_exon] + (run_k == m[run_exon] - 1) * rem[run_exon]
        words[ooff[:-1][run_exon] + 2 * run_k] = (run_len.astype(np.uint32) << np.uint32(4)) | np.uint32(M)
        ind_k = np.arange(len(ind_exon)) - np.repeat(np.cumsum(nind) - nind, nind)
        words[ooff[:-1][ind_exon] + 2 * ind_k + 1] = (ind_len.astype(np.uint32) << np.uint32(4)) | np.where(ind_isD, D, I).astype(np.uint32)
        nn = ~is_last
        words[ooff[1:][nn] - 1] = (gap[nn].astype(np.uint32) << np.uint32(4)) | np.uint32(N)
        ops_per_read = np.add.reduceat(per_exon_ops, eoff[:-1])
        qlen_core = np.add.reduceat(mb + isum, eoff[:-1])
        indel_bases = np.add.reduceat(dsum + isum, eoff[:-1])

    # low-coverage rejects: a soft clip longer than a third of the read
    lowcov = cls == 3
    lclip = np.where(lowcov, qlen_core, lclip)
    # add clips
    has_l = lclip > 0; has_r = rclip > 0
    tot_ops = ops_per_read + has_l + has_r
    coff = np.zeros(n + 1, np.int64); np.cumsum(tot_ops, out=coff[1:])
    cigar = np.zeros(coff[-1], np.uint32)
    core_start = coff[:-1] + has_l
    src_read = np.repeat(np.arange(n), ops_per_read)
    roff = np.zeros(n + 1, np.int64); np.cumsum(ops_per_read, out=roff[1:])
    dst = core_start[src_read] + (np.arange(roff[-1]) - roff[:-1][src_read])
    cigar[dst] = words
    cigar[coff[:-1][has_l]] = (lclip[has_l].astype(np.uint32) << np.uint32(4)) | np.uint32(S)
    cigar[coff[1:][has_r] - 1] = (rclip[has_r].astype(np.uint32) << np.uint32(4)) | np.uint32(S)

    qlen = qlen_core + lclip + rclip
    nm = (rng.random(n) * 0.05 * qlen_core).astype(np.int64) + indel_bases
    lowid = cls == 2
    nm = np.where(lowid, (qlen_core * 0.35).astype(np.int64) + indel_bases, nm)

    r = Reads()
    r.tid = tid.astype(np.int32)
    r.pos = (s[np.arange(n), np.argmax(keep, axis=1)] - 1).astype(np.int32)
    r.flag = np.where(rev, 16, 0).astype(np.uint16)
    r.l_qseq = qlen.astype(np.int32)
    r.nm = nm.astype(np.int32)
    xs = np.zeros(n, np.int8)
    tagged = rng.random(n) < 0.10
    xs[tagged] = np.where(rng.random(int(tagged.sum())) < 0.5, ord("+"), ord("-")).astype(np.int8)
    r.xs = xs
    r.qid = np.arange(n, dtype=np.int64) + QID_BASE[0]
    r.qname_hash = splitmix64(r.qid)
    r.cigar = cigar; r.cigar_off = coff.astype(np.uint64)

    if quirk_frac > 0:
        _inject_quirks(rng, r, quirk_frac)
    return r


def _inject_quirks(rng, r: Reads, frac: float):
    """Rewrite a few CIGARs in place (same op count) to hit the walk's corner cases: D>50 cuts, N<3 non-cuts,
    =/X ops, short internal exons.  Only ops of reads with >= 5 ops are touched, lengths stay consistent enough
    (l_qseq / NM are recomputed for '=' / 'X'; D and N do not consume query)."""
    n = r.n
    pick = np.nonzero((rng.random(n) < frac) & (np.diff(r.cigar_off.astype(np.int64)) >= 5))[0]
    for i in pick:
        lo, hi = int(r.cigar_off[i]), int(r.cigar_off[i + 1])
        c = r.cigar[lo:hi]
        kind = int(rng.integers(0, 5))
        mid = [k for k in range(1, len(c) - 1) if (c[k] & 15) == M]
        nn = [k for k in range(1, len(c) - 1) if (c[k] & 15) == N]
        if kind == 0 and nn:          # a long deletion instead of an intron (cuts when > max_delet)
            k = nn[0]; c[k] = (np.uint32(int(rng.integers(45, 60))) << np.uint32(4)) | np.uint32(D)
        elif kind == 1 and nn:        # a 1-2 bp N (does not cut)
            k = nn[-1]; c[k] = (np.uint32(int(rng.integers(1, 3))) << np.uint32(4)) | np.uint32(N)
        elif kind == 2 and mid:       # '=' and 'X' ops
            for k in mid[:2]:
                c[k] = (c[k] & ~np.uint32(15)) | np.uint32(EQ if rng.random() < 0.5 else X)
        elif kind == 3 and mid:       # 1-2 bp internal exon (vanishes, bam2gtf.c:45)
            k = mid[len(mid) // 2]
            old = int(c[k] >> 4); new = int(rng.integers(1, 3))
            c[k] = (np.uint32(new) << np.uint32(4)) | np.uint32(M)
            r.l_qseq[i] -= old - new
        elif kind == 4 and mid:       # exactly 3 bp internal exon (kept) / deletion of exactly 50 (no cut)
            k = mid[0]
            old = int(c[k] >> 4)
            c[k] = (np.uint32(3) << np.uint32(4)) | np.uint32(M)
            r.l_qseq[i] -= old - 3
            if nn:
                c[nn[0]] = (np.uint32(50) << np.uint32(4)) | np.uint32(D)
        r.cigar[lo:hi] = c


def make_sj(reads_soa_exons, frac: float = 0.7, seed: int = 5):
    """SJ.out.tab rows from read-derived junctions: `frac` of the distinct (tid, don, acc) with uniq U[0,5], multi U[0,2].
    reads_soa_exons = (tid per read, exon_off, exon_start, exon_end) of the read chains (from the oracle's CIGAR walk)."""
    rng = np.random.default_rng(seed)
    tid, off, es, ee = reads_soa_exons
    ne = np.diff(off.astype(np.int64))
    rid = np.repeat(np.arange(len(tid)), ne)
    last = np.ones(len(es), bool); last[off[1:].astype(np.int64) - 1] = False   # True where a junction follows
    j = np.nonzero(last)[0]
    don = ee[j].astype(np.int64) + 1; acc = es[j + 1].astype(np.int64) - 1; jt = tid[rid[j]].astype(np.int64)
    key = np.unique(np.stack([jt, don, acc], 1), axis=0)
    keepm = rng.random(len(key)) < frac
    key = key[keepm]
    uniq = rng.integers(0, 6, len(key)); multi = rng.integers(0, 3, len(key))
    return dict(tid=key[:, 0].astype(np.int32), don=key[:, 1].astype(np.int32), acc=key[:, 2].astype(np.int32),
                uniq_c=uniq.astype(np.int32), multi_c=multi.astype(np.int32))


def read_junctions(reads: Reads, min_intron: int = 3):
    """(tid, don, acc) of every N op of at least min_intron bases, STAR's SJ.out.tab convention (first / last intron base, 1-based):
    what bam2gtf's cuts leave as exon[j].end + 1 / exon[j+1].start - 1, up to its rare corner cases (vanishing short exons, D cuts)."""
    off = reads.cigar_off.astype(np.int64)
    cig = reads.cigar
    op = cig & np.uint32(15); ln = (cig >> np.uint32(4)).astype(np.int64)
    consume = np.where((op == M) | (op == D) | (op == N) | (op == EQ) | (op == X), ln, 0)
    cs = np.zeros(len(cig) + 1, np.int64); np.cumsum(consume, out=cs[1:])
    rid = np.repeat(np.arange(reads.n), np.diff(off))
    before = cs[:-1] - cs[off[:-1]][rid]                      # reference bases consumed by the read's earlier ops
    j = np.nonzero((op == N) & (ln >= min_intron))[0]
    don = reads.pos[rid[j]].astype(np.int64) + before[j] + 1
    return reads.tid[rid[j]].astype(np.int64), don, don + ln[j] - 1


def make_sj_from_reads(reads: Reads, frac: float = 0.7, seed: int = 5, min_intron: int = 3):
    """SJ.out.tab rows: `frac` of the distinct read-derived junctions with uniq U[0,5], multi U[0,2], sorted by (tid, don, acc)."""
    rng = np.random.default_rng(seed)
    t, d, a = read_junctions(reads, min_intron)
    key = np.unique((t << 52) | (d << 26) | a) if len(t) and d.max() < (1 << 26) and a.max() < (1 << 26) else None
    if key is not None:
        kt, kd, ka = key >> 52, (key >> 26) & ((1 << 26) - 1), key & ((1 << 26) - 1)
    else:
        k3 = np.unique(np.stack([t, d, a], 1), axis=0) if len(t) else np.zeros((0, 3), np.int64)
        kt, kd, ka = k3[:, 0], k3[:, 1], k3[:, 2]
    keepm = rng.random(len(kt)) < frac
    kt, kd, ka = kt[keepm], kd[keepm], ka[keepm]
    return dict(tid=kt.astype(np.int32), don=kd.astype(np.int32), acc=ka.astype(np.int32),
                uniq_c=rng.integers(0, 6, len(kt)).astype(np.int32), multi_c=rng.integers(0, 3, len(kt)).astype(np.int32))


def concat_reads(parts):
    """Batches of consecutive chromosome blocks (each coordinate sorted) as one stream."""
    r = Reads()
    for k in ("tid", "pos", "flag", "l_qseq", "nm", "xs", "qname_hash", "qid"):
        setattr(r, k, np.concatenate([getattr(p, k) for p in parts]))
    r.cigar = np.concatenate([p.cigar for p in parts])
    base = np.cumsum([0] + [int(p.cigar_off[-1]) for p in parts])
    r.cigar_off = np.concatenate([parts[i].cigar_off[:-1].astype(np.int64) + base[i] for i in range(len(parts))] + [base[-1:]]).astype(np.uint64)
    r.chrom_names, r.chrom_lens = parts[0].chrom_names, parts[0].chrom_lens
    return r


def clone_chromosome(anno_soa: dict, reads: Reads, sj: dict, n_chrom: int):
    """The same annotation / reads / junctions once more on `n_chrom` further chromosomes (tid + n_chrom) at the SAME coordinates, with
    gene ids of their own: split pieces of the copy meet equal chains on the original chromosome (the cross-locus case)."""
    a = dict(anno_soa)
    ng = int(a["gene"].max()) + 1 if len(a["gene"]) else 0
    a2 = {k: np.concatenate([a[k], a[k]]) for k in ("start", "end", "is_rev", "exon_start", "exon_end")}
    a2["tid"] = np.concatenate([a["tid"], a["tid"] + n_chrom]).astype(np.int32)
    a2["gene"] = np.concatenate([a["gene"], a["gene"] + ng]).astype(np.int32)
    off = a["exon_off"].astype(np.int64)
    a2["exon_off"] = np.concatenate([off[:-1], off + off[-1]]).astype(np.uint32)
    r2 = reads.take(np.arange(reads.n))
    r2.tid = (r2.tid + n_chrom).astype(np.int32)
    r2.qid = r2.qid + (int(reads.qid.max()) + 1 if reads.n else 0)
    r2.qname_hash = splitmix64(r2.qid)
    both = _concat(reads, r2)
    both.chrom_names = [f"chr{i + 1}" for i in range(2 * n_chrom)]; both.chrom_lens = list(reads.chrom_lens[:n_chrom]) * 2 if reads.chrom_lens else None
    s2 = {k: np.concatenate([sj[k], sj[k]]) for k in ("don", "acc", "uniq_c", "multi_c")}
    s2["tid"] = np.concatenate([sj["tid"], sj["tid"] + n_chrom]).astype(np.int32)
    return a2, both, s2


# ----------------------------------------------------------------------------- text renderers (reference inputs)

def cigar_string(words) -> str:
    return "".join(f"{int(w) >> 4}{OPS[int(w) & 15]}" for w in words)


def write_sam(path, reads: Reads, with_seq: bool = True, shuffle_tags: bool = False, nh=None):
    """nh: optional per-record NH:i values (0 = no tag) for bam2sj inputs."""
    with open(path, "w") as f:
        f.write("@HD\tVN:1.0\tSO:coordinate\n")
        for nme, ln in zip(reads.chrom_names, reads.chrom_lens):
            f.write(f"@SQ\tSN:{nme}\tLN:{ln}\n")
        co = reads.cigar_off
        for i in range(reads.n):
            cg = cigar_string(reads.cigar[co[i]:co[i + 1]])
            seq = "A" * int(reads.l_qseq[i]) if with_seq else "*"
            tags = f"NM:i:{int(reads.nm[i])}"
            if reads.xs[i]:
                tags += f"\tXS:A:{chr(int(reads.xs[i]))}"
            if nh is not None and nh[i]:
                tags += f"\tNH:i:{int(nh[i])}"
            chrom = reads.chrom_names[int(reads.tid[i])] if reads.tid[i] >= 0 else "*"
            f.write(f"read{int(reads.qid[i])}\t{int(reads.flag[i])}\t{chrom}\t{int(reads.pos[i]) + 1}\t60\t{cg}\t*\t0\t0\t{seq}\t*\t{tags}\n")


def write_gtf(path, anno: Annotation, rows=None):
    with open(path, "w") as f:
        rng_rows = range(anno.n_trans) if rows is None else rows
        for t in rng_rows:
            chrom = anno.chrom_names[int(anno.tid[t])]
            strand = "-" if anno.is_rev[t] else "+"
            g = int(anno.gene[t])
            attr = f'gene_id "G{g}"; transcript_id "G{g}.T{t}"; gene_name "GN{g}"; transcript_name "G{g}.TN{t}";'
            f.write(f"{chrom}\tsynth\ttranscript\t{int(anno.start[t])}\t{int(anno.end[t])}\t.\t{strand}\t.\t{attr}\n")
            lo, hi = int(anno.exon_off[t]), int(anno.exon_off[t + 1])
            order = range(hi - 1, lo - 1, -1) if anno.is_rev[t] else range(lo, hi)
            for k in order:
                f.write(f"{chrom}\tsynth\texon\t{int(anno.exon_start[k])}\t{int(anno.exon_end[k])}\t.\t{strand}\t.\t{attr}\n")


def write_rm_gtf(path, rrna, chrom_names):
    with open(path, "w") as f:
        for i in range(len(rrna["tid"])):
            chrom = chrom_names[int(rrna["tid"][i])]
            attr = f'gene_id "RR{i}"; transcript_id "RR{i}.1"; gene_name "RRN{i}"; transcript_name "RRN{i}.1";'
            f.write(f"{chrom}\tsynth\texon\t{int(rrna['start'][i])}\t{int(rrna['end'][i])}\t.\t+\t.\t{attr}\n")


def write_sj(path, sj, chrom_names, seed: int = 7):
    rng = np.random.default_rng(seed)
    order = rng.permutation(len(sj["tid"]))   # file order is irrelevant: the reader sorts (gtf.c:447)
    with open(path, "w") as f:
        for i in order:
            f.write(f"{chrom_names[int(sj['tid'][i])]}\t{int(sj['don'][i])}\t{int(sj['acc'][i])}\t1\t1\t0\t{int(sj['uniq_c'][i])}\t{int(sj['multi_c'][i])}\t30\n")
