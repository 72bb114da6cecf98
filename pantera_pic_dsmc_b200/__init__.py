"""B200-native particle hot path for PANTERA PIC-DSMC (move + sort + collide + deposit/gather).

The product is ``csrc/libpantera_gpu.so`` (hand-written CUDA for sm_100a behind the C ABI of
``include/pantera_gpu.h``); this package is the thin host-side harness around it: the ctypes
binding (:mod:`capi`), mesh/data-file fixtures (:mod:`meshgen`, :mod:`datafiles`) and the
synthetic workloads used by tests and ``bench.py`` (:mod:`workloads`).
"""
from . import capi, datafiles, meshgen, workloads  # noqa: F401

__all__ = ["capi", "datafiles", "meshgen", "workloads"]
