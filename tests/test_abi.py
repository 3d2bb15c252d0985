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
