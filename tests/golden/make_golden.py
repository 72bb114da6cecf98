#!/usr/bin/env python
"""Generates the fixtures under tests/golden/ (run from the repo root: `python tests/golden/make_golden.py`).

What these fixtures are -- and are not.  The reference (Fortran + MPI + PETSc) ships no golden vectors and cannot
be built in this image, so none of these files is reference OUTPUT.  They are:

  mt19937_64.json   known answers of the reference's generator (src/mt19937.f90 is the standard MT19937-64):
                    produced by libstdc++'s std::mt19937_64, an implementation independent of the oracle, for the
                    seeds the reference derives (RNG_SEED_GLOBAL*(PROC_ID+1), src/initialization.f90:2294-2296,
                    example seed 65644 from the reference's `input` file) and mapped through genrand64_real1.
  vss_tables.json   VSS pair tables of data/ar.* and data/air.* evaluated with the formulas of
                    src/initialization.f90:675-698 in plain Python (math.gamma), independent of datafiles.py.
  *.npz             SNAPSHOTS of the CPU oracle on small seeded inputs (inputs and outputs both stored).  They pin
                    the oracle against accidental change and give the GPU tests a fixture that does not execute
                    the oracle; their authority is the oracle's, i.e. "parity unpinned" by the reference.
"""
import json
import math
import os
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from pantera_pic_dsmc_b200 import capi, datafiles as df, workloads as wl  # noqa: E402
from oracle.binding import load_oracle  # noqa: E402

CXX = r"""
#include <cstdint>
#include <cstdio>
#include <random>
int main() {
    const unsigned long long seeds[] = {5489ULL, 65644ULL, 2ULL * 65644ULL, 8ULL * 65644ULL, 1ULL};
    std::printf("{");
    for (int s = 0; s < 5; s++) {
        std::mt19937_64 g(seeds[s]);
        std::printf("%s\"%llu\": {\"first\": [", s ? ", " : "", seeds[s]);
        unsigned long long x = 0;
        for (int i = 1; i <= 10000; i++) {
            x = g();
            if (i <= 8) std::printf("%s%llu", i > 1 ? ", " : "", x);
        }
        std::printf("], \"n10000\": %llu}", x);
    }
    std::printf("}\n");
    return 0;
}
"""


def make_mt():
    with tempfile.TemporaryDirectory() as t:
        src, exe = os.path.join(t, "kat.cpp"), os.path.join(t, "kat")
        open(src, "w").write(CXX)
        subprocess.check_call(["g++", "-O1", "-o", exe, src])
        kat = json.loads(subprocess.check_output([exe]).decode())
    for v in kat.values():  # genrand64_real1: (x >> 11) / (2^53 - 1), src/mt19937.f90:220-225
        v["rf_first"] = [float((x >> 11) * (1.0 / 9007199254740991.0)).hex() for x in v["first"]]
    json.dump(kat, open(os.path.join(HERE, "mt19937_64.json"), "w"), indent=1)


def make_vss():
    """sigma = pi (0.5 (d_i + d_j))^2; omega, alpha arithmetic means; CREF = sqrt(2 k Tref / m_r) *
    Gamma(5/2 - omega)^(-1/(2 (omega - 1/2)))  (or * 1.2354 when omega = 1/2), src/initialization.f90:675-698."""
    out = {}
    kb = 1.380649e-23
    for kind in ("ar", "air"):
        rows = {}
        for line in open(os.path.join(ROOT, "tests", "data", f"{kind}.species")):
            t = line.split("!")[0].split()
            if len(t) >= 8:
                rows[t[0]] = dict(mass=float(t[2]))   # NAME MOLWT MOLECULAR_MASS ... (src/initialization.f90:560-600)
        for line in open(os.path.join(ROOT, "tests", "data", f"{kind}.vss")):
            t = line.split("!")[0].split()
            if len(t) >= 5 and t[0] in rows:
                rows[t[0]].update(diam=float(t[1]), omega=float(t[2]), tref=float(t[3]), alpha=float(t[4]))
        names = [n for n, r in rows.items() if "diam" in r]
        tab = {}
        for a in names:
            for b in names:
                ra, rb = rows[a], rows[b]
                om = 0.5 * (ra["omega"] + rb["omega"])
                tref = 0.5 * (ra["tref"] + rb["tref"])
                mr = ra["mass"] * rb["mass"] / (ra["mass"] + rb["mass"])
                cref = math.sqrt(2 * kb * tref / mr)
                # `1.2354` is a default-REAL (single precision) literal in the Fortran source (:691): 1.23539996...
                cref *= float(np.float32(1.2354)) if abs(om - 0.5) < 1e-6 else math.gamma(2.5 - om) ** (-0.5 / (om - 0.5))
                tab[f"{a}|{b}"] = dict(sigma=math.pi * (0.5 * (ra["diam"] + rb["diam"])) ** 2, omega=om,
                                       alpha=0.5 * (ra["alpha"] + rb["alpha"]), cref=cref)
        out[kind] = tab
    json.dump(out, open(os.path.join(HERE, "vss_tables.json"), "w"), indent=1)


# ---------------------------------------------------------------------------------------------- oracle snapshots
CASES = {
    # name: (constructor, kwargs, config overrides, call, rng_mode)
    "advect_c5_tets": ("c5_channel", dict(ncx=4, ncy=3, ncz=3, ppc=12, jitter=0.15, dt_frac=0.6),
                       dict(collision_type=capi.NO_COLL), "advect", 1),
    "advect_c3_tris_walls": ("c3_tri_box", dict(nx=6, ny=5, n=3000, dt=4e-6), {}, "advect", 1),
    "advect_c3_axi": ("c3_tri_box", dict(nx=6, ny=5, n=3000, dt=4e-6, axi=True, lo=(0.0, 0.01), hi=(0.1, 0.06)), {},
                      "advect", 1),
    "advect_c1_periodic": ("c1_argon_box", dict(n=3000, nx=7, ny=5), dict(collision_type=capi.NO_COLL, dt=3e-4),
                           "advect", 1),
    "mcc_vss_air": ("mcc_box", dict(kind="air", n=3000, T=15000.0, bg_dens=5e21), dict(capacity=20000),
                    "mcc_collisions", 1),
    "mcc_vahedi_ccp": ("c4_ccp_1d", dict(ncells=16, n_per_species=1500), dict(dt=2e-9, capacity=20000), "mcc_vahedi", 1),
    "deposit_c4_line": ("c4_ccp_1d", dict(ncells=32, n_per_species=1500), {}, "deposit", 1),
    "collide_c1_mt": ("c1_argon_box", dict(n=4000, nx=3, ny=3), {}, "collide", 0),
    "collide_c2_air_philox": ("c2_air_box", dict(n=4000, nx=2, ny=2), {}, "collide", 1),
}


def build_case(name):
    ctor, kw, over, call, mode = CASES[name]
    w = getattr(wl, ctor)(**kw)
    for k, v in over.items():
        setattr(w.cfg, k, v)
    if "dt" in over:
        w.particles["DTRIM"] = w.cfg.dt
    w.cfg.track_id = 1
    w.particles["ID"] = np.arange(len(w.particles), dtype=np.int64) << 8
    return w, call, mode


def run_case(lib, w, call):
    """Returns (particles by ID, extras)."""
    wl.apply(lib, w)
    extra = {}
    if call == "advect":
        lib.advect(3)
    elif call == "collide":
        extra["counts"] = np.array(lib.collide(3), dtype=np.int64)
    elif call == "mcc_collisions":
        extra["counts"] = np.array(lib.mcc_collisions(3), dtype=np.int64)
    elif call == "mcc_vahedi":
        extra["counts"] = np.array(lib.mcc_vahedi(3), dtype=np.int64)
    elif call == "deposit":
        nn = w.mesh.nn
        extra["rhs"] = np.asarray(lib.deposit_charge(nn))
        if hasattr(lib, "deposit_charge_compensated") and lib.prefix == "po_":
            ref, ab = lib.deposit_charge_compensated(nn)
            extra["rhs_compensated"], extra["rhs_abs"] = np.asarray(ref), np.asarray(ab)
    return lib.download_particles(), extra


def make_snapshots():
    for name in CASES:
        w, call, mode = build_case(name)
        lib = load_oracle(rng_mode=mode)
        out, extra = run_case(lib, w, call)
        lib.close()
        np.savez_compressed(os.path.join(HERE, name + ".npz"), inp=w.particles, out=out, **extra)
        print(f"{name}: {len(w.particles)} -> {len(out)} particles", {k: v.shape for k, v in extra.items()})


if __name__ == "__main__":
    make_mt()
    make_vss()
    make_snapshots()
