"""The C-ABI library loads without a GPU, exports every entry point include/lr2rmats_b200.h declares, and fails loudly
(no CPU fallback) when no device is present."""
import ctypes
import os
import re

import pytest

from lr2rmats_b200 import api, cabi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_functions():
    src = open(os.path.join(ROOT, "include", "lr2rmats_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(lrb_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    if not os.path.exists(api.LIB_PATH):
        import __graft_entry__
        __graft_entry__.build()
    lib = ctypes.CDLL(api.LIB_PATH)
    names = header_functions()
    assert len(names) >= 25
    for n in names:
        assert hasattr(lib, n), f"{n} declared in the header but not exported"
    assert set(api.ENTRY_POINTS) <= set(names)


def test_struct_sizes_match_header():
    # spot checks of the ctypes mirrors against the C layout (LP64)
    assert ctypes.sizeof(cabi.Batch) == 8 + 9 * 8
    assert ctypes.sizeof(cabi.FilterParams) == 16
    assert ctypes.sizeof(cabi.ExonParams) == 12
    assert ctypes.sizeof(cabi.UpdateParams) == 36
    assert ctypes.sizeof(cabi.MergedList) == 8 + 7 * 8
    assert ctypes.sizeof(cabi.UpdateResult) == ctypes.sizeof(cabi.ExonResult) + 3 * 8 + 4 * 8 + ctypes.sizeof(cabi.TransList) + \
        ctypes.sizeof(cabi.MergedList) + 19 * 4 + 4 + ctypes.sizeof(cabi.BedList)


def test_no_cpu_fallback_without_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(api.LrbError) as e:
        api.Context(0)
    assert e.value.code == -6   # LRB_E_NODEVICE


def test_shard_cuts_host_helper():
    import numpy as np
    tid = np.array([0, 0, 0, 0, 1, 1, 1, 1], np.int32)
    start = np.array([10, 20, 100, 110, 5, 6, 50, 60], np.int32)
    end = np.array([30, 40, 120, 130, 20, 21, 70, 80], np.int32)
    cuts = api.shard_cuts(tid, start, end, 2)
    assert cuts.tolist() == [0, 4, 8]
    cuts = api.shard_cuts(tid, start, end, 4)
    assert cuts[0] == 0 and cuts[-1] == 8 and all(np.diff(cuts) >= 0)
    for c in cuts[1:-1]:
        if 0 < c < 8:   # every cut is at a locus gap
            assert tid[c] != tid[c - 1] or start[c] > end[:c][tid[:c] == tid[c]].max()


def test_shard_planning_is_locus_and_qname_safe():
    """multi.plan_shards / lrb_shard_cuts_weighted (host helpers, no GPU): every cut sits where the record starts beyond every earlier end
    on its chromosome, never inside a qname run, and the shards are balanced by CIGAR ops."""
    import numpy as np
    from lr2rmats_b200 import multi, synth
    anno = synth.make_annotation(400, n_chrom=3, seed=5)
    reads = synth.make_reads(anno, 20_000, seed=6, reject_frac=0.3)
    b = reads.soa()
    for n_sh in (2, 5, 8):
        cuts = multi.plan_shards(b, n_sh)
        assert cuts[0] == 0 and cuts[-1] == reads.n and np.all(np.diff(cuts) > 0)
        start, end = multi.ref_span(b)
        key_s = ((b["tid"].astype(np.int64) + 1) << 32) | start; key_e = ((b["tid"].astype(np.int64) + 1) << 32) | end
        run_max = np.maximum.accumulate(key_e)
        for c in cuts[1:-1]:
            assert key_s[c] > run_max[c - 1], "cut inside a locus"
            assert b["qname_hash"][c] != b["qname_hash"][c - 1], "cut inside a qname run"
        ops = np.diff(b["cigar_off"].astype(np.int64)) + 16
        per = np.array([ops[cuts[k]:cuts[k + 1]].sum() for k in range(n_sh)], float)
        assert per.max() / per.mean() < 1.25
        # the slices reassemble the stream
        parts = [multi.take_shard(b, int(cuts[k]), int(cuts[k + 1])) for k in range(n_sh)]
        assert sum(len(p["tid"]) for p in parts) == reads.n
        assert np.array_equal(np.concatenate([p["cigar"] for p in parts]), b["cigar"])
        assert all(p["cigar_off"][0] == 0 and p["cigar_off"][-1] == len(p["cigar"]) for p in parts)


def test_header_binds_from_plain_c(tmp_path):
    """The boundary is a C ABI: the header compiles as strict C99 and a C caller (what the reference's main.c would be) links against the
    library and gets LRB_E_NODEVICE here (no GPU, no CPU fallback)."""
    import subprocess
    from lr2rmats_b200 import api
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    src = tmp_path / "bind.c"
    src.write_text('#include "lr2rmats_b200.h"\n#include <stdio.h>\n'
                   'int main(void) { lrb_ctx *g = NULL; lrb_filter_params p = {0.67f, 0.75f, 0.98f, 0}; lrb_sj_params s = {3, 1}; (void)p; (void)s;\n'
                   '  int rc = lrb_ctx_create(0, &g); printf("%s %d\\n", lrb_version(), rc); if (rc == LRB_OK) lrb_ctx_destroy(g);\n'
                   '  return (rc == LRB_OK || rc == LRB_E_NODEVICE) ? 0 : 1; }\n')
    libdir = os.path.dirname(api.LIB_PATH)
    exe = tmp_path / "bind"
    subprocess.run(["gcc", "-std=c99", "-pedantic", "-Wall", "-Werror", "-I", os.path.join(root, "include"), str(src), "-L", libdir, "-llr2rmats_b200",
                    f"-Wl,-rpath,{libdir}", "-o", str(exe)], check=True)
    p = subprocess.run([str(exe)], capture_output=True, text=True)
    assert p.returncode == 0 and "lr2rmats_b200" in p.stdout
