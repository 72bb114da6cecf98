#!/usr/bin/env python
"""Times the two round-2 additions next to the hot path on one GPU and prints one JSON line each:

* pgpu_field_solve on a 2-D triangle mesh of C2's size (2e5 triangles, 1e5 nodes): iterations, ms per solve, ms per
  conjugate-gradient iteration (two kernels), cold start and warm start;
* pgpu_reassign_particles_to_cells for 4e6 particles on the same mesh.

usage: python tools/field_rebin_bench.py [nx ny nparticles]
"""
import json
import sys
import time

import numpy as np

sys.path.insert(0, __file__.rsplit("/", 2)[0])
from pantera_pic_dsmc_b200 import capi, workloads as wl  # noqa: E402


def main():
    nx, ny, n = (int(a) for a in sys.argv[1:4]) if len(sys.argv) >= 4 else (400, 250, 4_000_000)
    w = wl.c3_tri_box(nx=nx, ny=ny, n=n, lo=(0.0, 0.0), hi=(0.4, 0.25))
    w.bcs = dict(particle_bc=[capi.SPECULAR] * 4, field_bc=[capi.RF_VOLTAGE_BC, capi.DIRICHLET_BC, capi.NEUMANN_BC, capi.NO_BC],
                 wall_temp=[300.0] * 4, acc_n=[1.0] * 4, acc_t=[1.0] * 4)
    w.cfg.dt = 1e-9
    gpu = capi.load_gpu(0)
    wl.apply(gpu, w)
    gpu.set_field_bcs(wall_potential=[0.0, 0.0, 0.0, 0.0], wall_rf_potential=[200.0, 0.0, 0.0, 0.0], rf_frequency=[13.56e6, 0.0, 0.0, 0.0],
                      wall_efield=[0.0, 0.0, 100.0, 0.0])
    rhs = np.random.default_rng(1).standard_normal(w.mesh.nn)
    out = {"what": "pgpu_field_solve", "triangles": w.mesh.nc, "nodes": w.mesh.nn}
    import os
    for label, tol, coop in (("rtol_1e-5", 1e-5, "1"), ("rtol_1e-10", 1e-10, "1"), ("rtol_1e-5_kernel_per_half_iteration", 1e-5, "0")):
        os.environ["PANTERA_GPU_CG_COOP"] = coop
        gpu.set_field_bcs(wall_potential=[0.0] * 4, wall_rf_potential=[200.0, 0.0, 0.0, 0.0], rf_frequency=[13.56e6, 0.0, 0.0, 0.0],
                          wall_efield=[0.0, 0.0, 100.0, 0.0])          # (drops the warm start, rebuilds the matrix)
        t0 = time.perf_counter()
        _, _, it, rel = gpu.field_solve(1, w.mesh.nn, w.mesh.nc, rhs=rhs, rtol=tol, max_iter=20000)
        t1 = time.perf_counter()
        _, _, it2, rel2 = gpu.field_solve(2, w.mesh.nn, w.mesh.nc, rhs=rhs, rtol=tol, max_iter=20000)   # next timestep: warm start
        t2 = time.perf_counter()
        out[label] = {"cold": {"iters": it, "relres": rel, "ms_with_setup": 1e3 * (t1 - t0)},
                      "warm_next_step": {"iters": it2, "relres": rel2, "ms": 1e3 * (t2 - t1), "ms_per_iter": 1e3 * (t2 - t1) / max(it2, 1)}}
    print(json.dumps(out))
    rng = np.random.default_rng(2)
    w.particles["IC"] = rng.integers(1, w.mesh.nc + 1, len(w.particles))
    gpu.upload_particles(w.particles)
    t0 = time.perf_counter()
    removed = gpu.reassign_particles_to_cells()
    t1 = time.perf_counter()
    print(json.dumps({"what": "pgpu_reassign_particles_to_cells", "particles": len(w.particles), "cells": w.mesh.nc, "removed": removed,
                      "ms": 1e3 * (t1 - t0), "particles_per_s": len(w.particles) / (t1 - t0)}))
    gpu.close()


if __name__ == "__main__":
    main()
