"""lr2rmats_b200 -- B200-native implementation of the lr2rmats per-alignment hot path.

The product is the CUDA library `lr2rmats_b200/csrc/liblr2rmats_b200.so` behind the C ABI of
`include/lr2rmats_b200.h`, plus the drop-in CLI `lr2rmats_b200/host/lr2rmats-b200`.  This Python package only holds the
ctypes binding used by tests and bench.py (`api`), the struct mirrors (`cabi`) and the synthetic workload generator
(`synth`).  There is no CPU compute path here: `api.Library()` raises if the CUDA library or a GPU is missing.
"""
__all__ = ["api", "cabi", "synth"]
