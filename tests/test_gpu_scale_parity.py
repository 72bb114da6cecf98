"""Bit-exact ADVECT parity at the BENCH mesh size (parity tests stopped at 2e5 particles / 576 tets).

One step of 3e6 particles on the bench strip mesh (32x128x128 cubes x 6 tets = 3.1e6 tetrahedra, the mesh of
`bench.py` at N = 1), device vs oracle, with the warp chunk of the LEAN pass forced to its bench value (1024 slots;
PANTERA_GPU_ADVECT_CHUNK), so that the multi-block filter records, the refill loop and the hit list run at realistic
size.  All-specular walls: every particle is deterministic and must be bit-identical.  Bench walls (diffuse in y):
the wall-hit particles go through libm on shared Philox streams (rtol 1e-12), everything else stays bit-identical.
"""
import os

import numpy as np
import pytest

from pantera_pic_dsmc_b200 import capi, workloads as wl
from oracle.binding import load_oracle

import util

pytestmark = pytest.mark.gpu

MOVE_FIELDS = ("X", "Y", "Z", "VX", "VY", "VZ", "EROT", "EVIB")
N = 3_000_000


@pytest.fixture(scope="module")
def bench_strip():
    w = wl.c5_channel_strip(32, 128, 128, n=N, rank=0, world=1, seed=77)
    w.cfg.collision_type = capi.NO_COLL
    w.cfg.track_id = 1
    # a quarter of the particles fly 4 steps' worth: long hop chains across many cells
    rng = np.random.default_rng(5)
    fast = rng.random(N) < 0.25
    for f in ("VX", "VY", "VZ"):
        w.particles[f][fast] *= 4.0
    return w


@pytest.mark.timeout(900)
@pytest.mark.parametrize("walls", ["specular", "bench"])
def test_advect_bench_mesh_one_step(bench_strip, walls, monkeypatch):
    w = bench_strip
    monkeypatch.setenv("PANTERA_GPU_ADVECT_CHUNK", "1024")
    if walls == "specular":
        w.bcs["particle_bc"] = [capi.SPECULAR] * 6
    else:
        w.bcs["particle_bc"] = [capi.SPECULAR, capi.SPECULAR, capi.DIFFUSE, capi.DIFFUSE, capi.SPECULAR, capi.SPECULAR]
    gpu = capi.load_gpu(0)
    orc = load_oracle(rng_mode=1)
    try:
        wl.apply(gpu, w)
        wl.apply(orc, w)
        gpu.advect(1)
        orc.advect(1)
        a, b = util.by_id(orc.download_particles()), util.by_id(gpu.download_particles())
        assert len(a) == N and np.array_equal(a["ID"], b["ID"])
        assert np.array_equal(a["IC"], b["IC"]), "cell indices differ"
        moved = float(np.mean(a["IC"] != util.by_id(w.particles)["IC"]))
        assert moved > 0.5, f"only {moved:.2f} of the particles changed cell: not a crossing test"
        if walls == "specular":
            util.assert_particles_match(a, b, exact_fields=MOVE_FIELDS, what="bench mesh, specular walls")
        else:
            util.assert_particles_match(a, b, rtol=1e-12, what="bench mesh, bench walls")
            # a particle that met no diffuse wall kept its speed: those must be bit-identical
            w0 = util.by_id(w.particles)
            sp0 = w0["VX"] ** 2 + w0["VY"] ** 2 + w0["VZ"] ** 2
            spa = a["VX"] ** 2 + a["VY"] ** 2 + a["VZ"] ** 2
            free = np.abs(spa - sp0) <= 1e-9 * sp0
            assert free.mean() > 0.9
            for f in MOVE_FIELDS:
                assert np.array_equal(a[f][free], b[f][free]), f"{f} not bit-exact away from the diffuse walls"
    finally:
        gpu.close()
        orc.close()
