"""Helpers shared by the parity tests."""
import numpy as np

from pantera_pic_dsmc_b200 import capi, workloads as wl

FLOAT_FIELDS = ["X", "Y", "Z", "VX", "VY", "VZ", "EROT", "EVIB", "DTRIM"]
INT_FIELDS = ["IC", "S_ID"]


def by_id(p: np.ndarray) -> np.ndarray:
    return p[np.argsort(p["ID"], kind="stable")]


def setup_pair(gpu, orc, w: wl.Workload):
    """Push the same workload through both libraries (device keeps IDs so particles can be matched)."""
    w.cfg.track_id = 1
    wl.apply(orc, w)
    wl.apply(gpu, w)


def assert_particles_match(a: np.ndarray, b: np.ndarray, rtol=1e-12, atol=0.0, exact_fields=(), what=""):
    """a, b: particle arrays already in the same order.  Integer fields bit-exact, floating fields within
    rtol (north_star: 1e-12 relative in FP64); fields in `exact_fields` must be bit-identical."""
    assert len(a) == len(b), f"{what}: particle counts differ {len(a)} vs {len(b)}"
    for f in INT_FIELDS:
        bad = np.nonzero(a[f] != b[f])[0]
        assert bad.size == 0, f"{what}: {f} differs for {bad.size} particles, first {bad[:5]}: {a[f][bad[:5]]} vs {b[f][bad[:5]]}"
    for f in FLOAT_FIELDS:
        if f in exact_fields:
            bad = np.nonzero(a[f] != b[f])[0]
            assert bad.size == 0, f"{what}: {f} not bit-exact for {bad.size} particles, first {bad[:5]}"
        else:
            scale = np.maximum(np.abs(a[f]), np.abs(b[f]))
            err = np.abs(a[f] - b[f])
            tol = rtol * scale + atol
            bad = np.nonzero(err > tol)[0]
            assert bad.size == 0, (f"{what}: {f} differs beyond rtol={rtol} for {bad.size} particles; worst rel "
                                   f"{np.max(err[bad] / np.maximum(scale[bad], 1e-300)):.3e} at {bad[:5]}")


def exact_fraction(a: np.ndarray, b: np.ndarray, fields=("X", "Y", "Z", "VX", "VY", "VZ")) -> float:
    ok = np.ones(len(a), dtype=bool)
    for f in fields:
        ok &= a[f] == b[f]
    return float(ok.mean()) if len(a) else 1.0


def kinetic_energy(p, mass):
    m = np.asarray(mass)[p["S_ID"] - 1]
    return float(np.sum(0.5 * m * (p["VX"] ** 2 + p["VY"] ** 2 + p["VZ"] ** 2)))


def momentum(p, mass):
    m = np.asarray(mass)[p["S_ID"] - 1]
    return np.array([np.sum(m * p["VX"]), np.sum(m * p["VY"]), np.sum(m * p["VZ"])])
