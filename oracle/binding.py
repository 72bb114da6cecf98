"""ctypes binding of the CPU oracle (oracle/_build/liboracle.so, prefix ``po_``): TEST INFRASTRUCTURE ONLY.

Lives outside the product package on purpose: pantera_pic_dsmc_b200 never imports, loads or names the oracle; only tests/,
__graft_entry__.smoke() and the cpu_baseline / --impl reference legs of bench.py import this module.  The entry points
mirror include/pantera_gpu.h one for one (tests/test_abi.py), so the generic ``capi.Lib`` methods drive both libraries.
"""
import ctypes as C
import os

import numpy as np

from pantera_pic_dsmc_b200.capi import Lib, _f64, _p

ORACLE_LIB_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_build", "liboracle.so")


class OracleLib(Lib):
    """CPU oracle (test infrastructure). rng_mode 0 = MT19937-64 (reference), 1 = Philox (device)."""

    def __init__(self, rng_mode: int = 0, path: str = ORACLE_LIB_PATH):
        super().__init__(path, "po_", 0)
        self._call("set_rng_mode", C.c_int(rng_mode))

    def set_rng_mode(self, mode):
        self._call("set_rng_mode", C.c_int(mode))

    def rf(self) -> float:
        f = self.lib.po_rf
        f.restype = C.c_double
        return f(self.ctx)

    def pick(self, n: int) -> int:
        f = self.lib.po_pick
        f.restype = C.c_int
        return f(self.ctx, C.c_int(n))

    def rng_stream(self, kind, tID, e0, e1=0):
        self.lib.po_rng_stream(self.ctx, C.c_int(kind), C.c_int(tID), C.c_uint64(e0), C.c_uint64(e1))

    def mt_next_u64(self) -> int:
        f = self.lib.po_mt_next_u64
        f.restype = C.c_uint64
        return f(self.ctx)

    def philox(self, ctr, key):
        c = (C.c_uint32 * 4)(*ctr)
        k = (C.c_uint32 * 2)(*key)
        o = (C.c_uint32 * 4)()
        self.lib.po_philox4x32_10(c, k, o)
        return list(o)

    def maxwell(self, u, t, m):
        v = (C.c_double * 3)()
        self.lib.po_maxwell(self.ctx, *[C.c_double(x) for x in u], *[C.c_double(x) for x in t], C.c_double(m), v)
        return list(v)

    def internal_energy(self, dof, temp):
        f = self.lib.po_internal_energy
        f.restype = C.c_double
        return f(self.ctx, C.c_int(dof), C.c_double(temp))

    def flx(self, sn, tinf, m):
        f = self.lib.po_flx
        f.restype = C.c_double
        return f(self.ctx, C.c_double(sn), C.c_double(tinf), C.c_double(m))

    def interp_cs(self, e, te, tcs):
        f = self.lib.po_interp_cs
        f.restype = C.c_double
        te, tcs = _f64(te), _f64(tcs)
        return f(C.c_double(e), C.c_int(len(te)), _p(te), _p(tcs))

    def deposit_charge_compensated(self, nnodes):
        rhs, ab = np.zeros(nnodes), np.zeros(nnodes)
        self._call("deposit_charge_compensated", _p(rhs), _p(ab), C.c_int64(nnodes))
        return rhs, ab

    def get_timings(self):
        s = np.zeros(3)
        self._call("get_timings", _p(s), C.c_int32(3))
        return s


def load_oracle(rng_mode: int = 0) -> OracleLib:
    return OracleLib(rng_mode)
