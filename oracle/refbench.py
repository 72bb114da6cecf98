"""`bench.py --impl reference`: the reference's CPU execution of the hot path on the box's host cores.

The PANTERA binary (Fortran + MPI + PETSc) cannot be built in this image (no Fortran compiler, no MPI), so
what is timed is the oracle -- the C restatement of the reference algorithm -- in the reference's own
parallel model: one rank per x-strip (ASSIGN_CELLS_TO_PROCS "stripsx"), per step ADVECT -> EXCHANGE ->
DSMC_COLLISIONS, one host thread per rank (oracle/po_mpi_sim.c), MT19937-64 stream per rank seeded as
src/initialization.f90:2294-2296.  Test / benchmark infrastructure: lives under oracle/, outside the product package.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from pantera_pic_dsmc_b200 import capi, workloads as wl
from oracle.binding import load_oracle


def usable_cores() -> int:
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return max(1, os.cpu_count() or 1)


def make_ranks(nranks: int, cubes=(8, 32, 32), ppc: int = 40, seed: int = 7, u_drift: float = 100.0):
    """nranks oracle contexts, each seeded with its own strip of the C5 channel (same cell size, particles
    per cell, dt and wall models as the GPU arm's workload)."""
    ncx, ncy, ncz = cubes
    n = ncx * ncy * ncz * 6 * ppc
    libs, mesh = [], None
    for r in range(nranks):
        w = wl.c5_channel_strip(ncx, ncy, ncz, n=n, rank=r, world=nranks, seed=seed + r, mesh=mesh, u_drift=u_drift)
        mesh = w.mesh
        o = load_oracle(rng_mode=0)
        wl.apply(o, w)
        libs.append(o)
    return libs, n, mesh


def sim_run(libs, tID0: int, nsteps: int):
    lib = libs[0].lib
    f = lib.po_sim_run
    f.restype = C.c_int
    arr = (C.c_void_p * len(libs))(*[l.ctx for l in libs])
    sec = C.c_double()
    tot = (C.c_int64 * 3)()
    rc = f(arr, C.c_int(len(libs)), C.c_int(tID0), C.c_int(nsteps), C.byref(sec), tot)
    if rc != 0:
        raise capi.PanteraError(f"po_sim_run -> {rc}")
    return sec.value, list(tot)


def run(steps: int = 10, warmup: int = 3, world: int = 1, cubes=(8, 32, 32), ppc: int = 40):
    # 2.0e6 particles (189 MB of AoS records) per strip: like the real workload, a strip does not fit the caches
    # (a 2.5e5-particle strip runs 1.8x faster per particle on this build container's cores)
    cores = usable_cores()
    libs, n_per_rank, mesh = make_ranks(cores, cubes, ppc, u_drift=0.0)  # same config as the GPU arm: no mean drift
    sim_run(libs, 1, max(warmup, 1))
    sec, tot = sim_run(libs, 1 + warmup, steps)
    n_total = sum(l.num_particles() for l in libs)
    for l in libs:
        l.close()
    return {
        "value": n_total * steps / sec, "ms_per_step": sec / steps * 1e3, "cores": cores,
        "sample": (f"C5 channel, {cores} x-strips of {cubes[0]}x{cubes[1]}x{cubes[2]} cubes x6 tets, {n_per_rank} particles "
                   f"per strip ({n_total} total), {steps} steps, one host thread per strip (reference MPI model), "
                   f"{tot[2]} particles exchanged, {tot[0]} collisions"),
        "config": {"workload": "C5 3-D unstructured argon channel (VSS/NTC, diffuse y-walls, specular x/z) -- bounded "
                               "sample on host cores: same cell size, 40 particles per tet, same dt as the GPU arm",
                   "impl": "oracle port of the reference algorithm (the Fortran+MPI+PETSc binary cannot be built here)",
                   "threads": cores},
    }
