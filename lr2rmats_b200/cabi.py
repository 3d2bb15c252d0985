"""ctypes mirror of include/lr2rmats_b200.h: struct layouts + numpy <-> struct helpers.

Shared by the product binding (lr2rmats_b200/api.py, the CUDA library) and by the test-only oracle binding
(tests/oracle_port.py); it contains no compute.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

i32p, u32p, u16p, u8p, i8p, u64p = (C.POINTER(t) for t in (C.c_int32, C.c_uint32, C.c_uint16, C.c_uint8, C.c_int8, C.c_uint64))


class Batch(C.Structure):
    _fields_ = [("n", C.c_int64), ("tid", i32p), ("pos", i32p), ("flag", u16p), ("l_qseq", i32p), ("nm", i32p),
                ("xs", i8p), ("qname_hash", u64p), ("cigar_off", u64p), ("cigar", u32p)]


class Anno(C.Structure):
    _fields_ = [("n_trans", C.c_int32), ("n_exon", C.c_int64), ("tid", i32p), ("start", i32p), ("end", i32p),
                ("is_rev", u8p), ("gene", i32p), ("exon_off", u32p), ("exon_start", i32p), ("exon_end", i32p)]


class Sj(C.Structure):
    _fields_ = [("n", C.c_int64), ("tid", i32p), ("don", i32p), ("acc", i32p), ("uniq_c", i32p), ("multi_c", i32p)]


class SjParams(C.Structure):
    _fields_ = [("min_intron", C.c_int32), ("pair_only", C.c_int32)]

    @classmethod
    def default(cls, **kw):
        p = cls(3, 1)
        for k, v in kw.items():
            setattr(p, k, v)
        return p


def sj_to_np(s: "Sj") -> dict:
    return {k: _arr(getattr(s, k), s.n, np.int32) for k in ("tid", "don", "acc", "uniq_c", "multi_c")}


class Chains(C.Structure):
    _fields_ = [("n", C.c_int64), ("tid", i32p), ("is_rev", u8p), ("exon_off", u32p), ("exon_start", i32p), ("exon_end", i32p)]


class FilterParams(C.Structure):
    _fields_ = [("cov_rate", C.c_float), ("map_qual", C.c_float), ("sec_rat", C.c_float), ("min_intron_n", C.c_int32)]

    @classmethod
    def default(cls, **kw):
        p = cls(0.67, 0.75, 0.98, 0)
        for k, v in kw.items():
            setattr(p, k, v)
        return p


class ExonParams(C.Structure):
    _fields_ = [("min_exon", C.c_int32), ("min_intron", C.c_int32), ("max_delet", C.c_int32)]

    @classmethod
    def default(cls, **kw):
        p = cls(3, 3, 50)
        for k, v in kw.items():
            setattr(p, k, v)
        return p


class UpdateParams(C.Structure):
    _fields_ = [("min_sj_cnt", C.c_int32), ("ss_dis", C.c_int32), ("end_dis", C.c_int32), ("full_level", C.c_int32),
                ("split_trans", C.c_int32), ("use_multi", C.c_int32), ("force_strand", C.c_int32),
                ("single_exon_ovlp_frac", C.c_float), ("want_summary", C.c_int32)]

    @classmethod
    def default(cls, **kw):
        p = cls(1, 0, 0x7FFFFFFF, 5, 0, 0, 0, 0.80, 1)
        for k, v in kw.items():
            setattr(p, k, v)
        return p


class FilterResult(C.Structure):
    _fields_ = [("n", C.c_int64), ("pass_", u8p), ("score", i32p), ("intron_n", i32p), ("n_keep", C.c_int64), ("keep_idx", u32p)]


class ExonResult(C.Structure):
    _fields_ = [("n_reads", C.c_int64), ("read_idx", u32p), ("tid", i32p), ("is_rev", u8p), ("exon_off", u32p),
                ("exon_start", i32p), ("exon_end", i32p)]


class TransList(C.Structure):
    _fields_ = [("n", C.c_int64), ("read", u32p), ("exon_lo", u32p), ("exon_n", u32p), ("piece", i32p)]


class MergedList(C.Structure):
    _fields_ = [("n", C.c_int64), ("cand", u32p), ("cov", i32p), ("t_tid", i32p), ("t_start", i32p), ("t_end", i32p),
                ("first_start", i32p), ("last_end", i32p)]


class BedList(C.Structure):
    _fields_ = [("n", C.c_int64), ("tid", i32p), ("start", i32p), ("end", i32p), ("score", i32p), ("type", u8p), ("is_rev", u8p)]


S_COUNT = 19
S_NAMES = ["anno_genes", "anno_trans", "upd_genes", "novel_trans", "novel_full", "novel_partial", "novel_exons", "novel_sites",
           "novel_junc", "known_trans", "known_genes", "uniq_known", "novel_bam", "novel_reliable", "uniq_reliable",
           "novel_unreliable", "uniq_unreliable", "unrecog", "uniq_unrecog"]


class UpdateResult(C.Structure):
    _fields_ = [("ex", ExonResult), ("cls", u32p), ("ref_anno", i32p), ("exon_flag", u8p),
                ("n_known", C.c_int64), ("known_idx", u32p), ("n_unrecog", C.c_int64), ("unrecog_idx", u32p),
                ("novel", TransList), ("updated", MergedList), ("summary", C.c_int32 * S_COUNT), ("bed", BedList)]


class TransTable(C.Structure):
    _fields_ = [("n", C.c_int64), ("name_idx", u32p), ("piece", i32p), ("t_tid", i32p), ("t_start", i32p), ("t_end", i32p), ("t_rev", u8p),
                ("e_tid", i32p), ("e_rev", u8p), ("cov", i32p), ("ref_anno", i32p), ("exon_off", u32p), ("exon_start", i32p), ("exon_end", i32p)]


class UniqueResult(C.Structure):
    _fields_ = [("ex", ExonResult), ("uniq", MergedList), ("n_shared", C.c_int64), ("shared_idx", u32p)]


# class-word / flag bits
C_KNOWN, C_KNOWN_SITE, C_UNRELIABLE, C_FULL, C_LFULL, C_RFULL, C_LNOTH, C_RNOTH, C_SJ_CHECKED = (1 << k for k in range(9))
F_NOVEL_EXON, F_NOVEL_DON, F_NOVEL_ACC, F_NOVEL_JUNC, F_UNRELIABLE = (1 << k for k in range(5))

_DT = {"tid": np.int32, "pos": np.int32, "flag": np.uint16, "l_qseq": np.int32, "nm": np.int32, "xs": np.int8,
       "qname_hash": np.uint64, "cigar_off": np.uint64, "cigar": np.uint32, "start": np.int32, "end": np.int32,
       "is_rev": np.uint8, "gene": np.int32, "exon_off": np.uint32, "exon_start": np.int32, "exon_end": np.int32,
       "don": np.int32, "acc": np.int32, "uniq_c": np.int32, "multi_c": np.int32}


def _ptr(a: np.ndarray, ctype):
    return a.ctypes.data_as(C.POINTER(ctype))


_CT = {np.int32: C.c_int32, np.uint32: C.c_uint32, np.uint16: C.c_uint16, np.uint8: C.c_uint8, np.int8: C.c_int8, np.uint64: C.c_uint64}


def _fill(struct, arrays: dict, keep: list):
    for name, _ in struct._fields_:
        if name in arrays:
            a = np.ascontiguousarray(arrays[name], dtype=_DT[name])
            keep.append(a)
            setattr(struct, name, _ptr(a, _CT[_DT[name]]))


def make_batch(soa: dict):
    """soa: dict with the lrb_batch arrays. Returns (struct, keepalive)."""
    keep = []
    b = Batch()
    _fill(b, soa, keep)
    b.n = len(soa["tid"])
    return b, keep


def make_anno(soa: dict):
    keep = []
    a = Anno()
    if "gene" not in soa:
        soa = dict(soa, gene=np.zeros(len(soa["tid"]), np.int32))
    if "is_rev" not in soa:
        soa = dict(soa, is_rev=np.zeros(len(soa["tid"]), np.uint8))
    if "exon_off" not in soa:   # remove table: one exon per entry
        n = len(soa["tid"])
        soa = dict(soa, exon_off=np.arange(n + 1, dtype=np.uint32), exon_start=soa["start"], exon_end=soa["end"])
    _fill(a, soa, keep)
    a.n_trans = len(soa["tid"])
    a.n_exon = len(soa["exon_start"])
    return a, keep


def make_sj(soa: dict):
    keep = []
    s = Sj()
    _fill(s, soa, keep)
    s.n = len(soa["tid"])
    return s, keep


def make_chains(soa: dict):
    keep = []
    c = Chains()
    _fill(c, soa, keep)
    c.n = len(soa["tid"])
    return c, keep


def _arr(ptr, n, dtype):
    n = int(n)
    if n == 0 or not ptr:
        return np.zeros(0, dtype)
    return np.ctypeslib.as_array(ptr, shape=(n,)).astype(dtype, copy=True)


def filter_to_np(r: FilterResult) -> dict:
    return dict(pass_=_arr(r.pass_, r.n, np.uint8), score=_arr(r.score, r.n, np.int32), intron_n=_arr(r.intron_n, r.n, np.int32),
                keep_idx=_arr(r.keep_idx, r.n_keep, np.uint32))


def exon_to_np(r: ExonResult) -> dict:
    n = int(r.n_reads)
    off = _arr(r.exon_off, n + 1, np.uint32)
    ne = int(off[-1]) if n else 0
    return dict(n_reads=n, read_idx=_arr(r.read_idx, n, np.uint32) if r.read_idx else None, tid=_arr(r.tid, n, np.int32),
                is_rev=_arr(r.is_rev, n, np.uint8), exon_off=off, exon_start=_arr(r.exon_start, ne, np.int32),
                exon_end=_arr(r.exon_end, ne, np.int32))


def merged_to_np(m: MergedList) -> dict:
    return {k: _arr(getattr(m, k), m.n, np.uint32 if k == "cand" else np.int32)
            for k in ("cand", "cov", "t_tid", "t_start", "t_end", "first_start", "last_end")}


def update_to_np(r: UpdateResult) -> dict:
    ex = exon_to_np(r.ex)
    n, ne = ex["n_reads"], len(ex["exon_start"])
    return dict(ex=ex, cls=_arr(r.cls, n, np.uint32), ref_anno=_arr(r.ref_anno, n, np.int32), exon_flag=_arr(r.exon_flag, ne, np.uint8),
                known_idx=_arr(r.known_idx, r.n_known, np.uint32), unrecog_idx=_arr(r.unrecog_idx, r.n_unrecog, np.uint32),
                novel={k: _arr(getattr(r.novel, k), r.novel.n, np.int32 if k == "piece" else np.uint32) for k in ("read", "exon_lo", "exon_n", "piece")},
                updated=merged_to_np(r.updated), summary=np.array(list(r.summary), np.int32),
                bed={k: _arr(getattr(r.bed, k), r.bed.n, np.uint8 if k in ("type", "is_rev") else np.int32)
                     for k in ("tid", "start", "end", "score", "type", "is_rev")})


def table_to_np(t: TransTable) -> dict:
    n = int(t.n)
    off = _arr(t.exon_off, n + 1, np.uint32)
    ne = int(off[-1]) if n else 0
    d = {k: _arr(getattr(t, k), n, np.uint8 if k in ("t_rev", "e_rev") else (np.uint32 if k == "name_idx" else np.int32))
         for k in ("name_idx", "piece", "t_tid", "t_start", "t_end", "t_rev", "e_tid", "e_rev", "cov", "ref_anno")}
    d.update(exon_off=off, exon_start=_arr(t.exon_start, ne, np.int32), exon_end=_arr(t.exon_end, ne, np.int32))
    return d


def bed_to_np(b: BedList) -> dict:
    return {k: _arr(getattr(b, k), b.n, np.uint8 if k in ("type", "is_rev") else np.int32) for k in ("tid", "start", "end", "score", "type", "is_rev")}


def unique_to_np(r: UniqueResult) -> dict:
    return dict(ex=exon_to_np(r.ex), uniq=merged_to_np(r.uniq), shared_idx=_arr(r.shared_idx, r.n_shared, np.uint32))


def exon_struct_from_np(ex: dict):
    """numpy exon dict -> ExonResult struct view (for orc_update / chains inputs)."""
    keep = []
    e = ExonResult()
    e.n_reads = ex["n_reads"]
    for name, dt, ct in (("tid", np.int32, C.c_int32), ("is_rev", np.uint8, C.c_uint8), ("exon_off", np.uint32, C.c_uint32),
                         ("exon_start", np.int32, C.c_int32), ("exon_end", np.int32, C.c_int32)):
        a = np.ascontiguousarray(ex[name], dtype=dt); keep.append(a); setattr(e, name, _ptr(a, ct))
    if ex.get("read_idx") is not None:
        a = np.ascontiguousarray(ex["read_idx"], dtype=np.uint32); keep.append(a); e.read_idx = _ptr(a, C.c_uint32)
    return e, keep
