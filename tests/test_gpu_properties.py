"""Size-independent properties of the hot path at bench-like size (2e7 particles, 7.9e5 tetrahedra), where the CPU
oracle is too slow to be the checker: conservation, sortedness, permutation checksums, idempotence of the sort and
time reversal of the deterministic flight.  (Parity proper, against the oracle, is in test_gpu_advect/collide.py.)"""
import numpy as np
import pytest

from pantera_pic_dsmc_b200 import capi, workloads as wl

import util

pytestmark = pytest.mark.gpu

N = 20_000_000
CUBES = (8, 128, 128)


@pytest.fixture(scope="module")
def big():
    w = wl.c5_channel_strip(*CUBES, n=N, rank=0, world=1, seed=77)
    w.cfg.track_id = 1
    # all walls specular: the flight is deterministic and energy is conserved exactly by every operation
    w.bcs["particle_bc"] = [capi.SPECULAR] * 6
    return w


def test_full_step_conserves_count_energy_and_ids(gpu, big):
    wl.apply(gpu, big)
    m = big.species.mass
    e0 = util.kinetic_energy(big.particles, m)
    id_sum, id_xor = int(big.particles["ID"].sum()), int(np.bitwise_xor.reduce(big.particles["ID"]))
    ncoll, nreac = gpu.run_steps(1, 3)
    q = gpu.download_particles()
    assert len(q) == N and nreac == 0
    assert ncoll > 0.3 * 0.5 * 0.2 * N * 3          # N nu dt / 2 per step with dt = 0.2/nu (equilibrium estimate)
    # permutation: the same multiset of IDs came back (sum and xor checksums)
    assert int(q["ID"].sum()) == id_sum and int(np.bitwise_xor.reduce(q["ID"])) == id_xor
    # elastic VSS collisions and specular walls conserve kinetic energy (pairwise rounding only)
    assert util.kinetic_energy(q, m) == pytest.approx(e0, rel=1e-11)
    assert np.all((q["IC"] >= 1) & (q["IC"] <= big.mesh.nc))
    # every particle is inside the tetrahedron it claims (barycentric test through the face planes)
    fn = np.asarray(big.mesh.face_normal)          # (3, 4, nc) outward normals
    cn = np.asarray(big.mesh.cell_nodes)
    xyz = np.asarray(big.mesh.node_coords)
    sel = np.random.default_rng(0).integers(0, N, 200_000)
    ic = q["IC"][sel] - 1
    pos = np.stack([q["X"][sel], q["Y"][sel], q["Z"][sel]])
    for f in range(4):
        pf = xyz[:, cn[f, ic] - 1]
        dist = np.sum((pf - pos) * fn[:, f, ic], axis=0)   # outward normal: inside <=> distance to every face >= 0
        assert dist.min() > -1e-12, (f, dist.min())


def test_sort_sortedness_counts_and_idempotence(gpu, big):
    wl.apply(gpu, big)
    gpu.advect(1)
    gpu.sort()
    a = gpu.download_particles()
    assert np.all(np.diff(a["IC"]) >= 0)
    assert np.array_equal(gpu.get_cell_counts(big.mesh.nc), np.bincount(a["IC"] - 1, minlength=big.mesh.nc))
    gpu.sort()                                        # sorting a sorted store changes nothing but in-cell order
    b = gpu.download_particles()
    assert np.array_equal(a["IC"], b["IC"])
    assert np.array_equal(np.sort(a["ID"]), np.sort(b["ID"]))
    ka, kb = np.lexsort((a["ID"], a["IC"])), np.lexsort((b["ID"], b["IC"]))
    for f in ("X", "VX", "S_ID"):
        assert np.array_equal(a[f][ka], b[f][kb])


def test_flight_time_reversal(gpu, big):
    """Specular walls, no collisions: advect, reverse every velocity, advect again -> every particle is back where
    it started (to rounding: each cell crossing re-evaluates x + v t) and back in its starting cell."""
    big.cfg.collision_type = capi.NO_COLL
    wl.apply(gpu, big)
    gpu.advect(1)
    p = gpu.download_particles()
    for f in ("VX", "VY", "VZ"):
        p[f] = -p[f]
    gpu.upload_particles(p)
    gpu.advect(2)
    q = util.by_id(gpu.download_particles())
    s = util.by_id(big.particles)
    big.cfg.collision_type = capi.DSMC
    L = float(big.cfg.xmax)
    for f in ("X", "Y", "Z"):
        assert np.max(np.abs(q[f] - s[f])) < 1e-12 * max(L, 1.0), f
    for f in ("VX", "VY", "VZ"):
        assert np.allclose(q[f], -s[f], rtol=1e-12, atol=1e-9), f
    assert np.mean(q["IC"] == s["IC"]) > 0.99999


def test_deterministic_mode_reproduces_a_run_bit_for_bit():
    """pgpu_set_deterministic: stable sort -> the same input and seed give the same particles, in the same order, run
    after run (diffuse walls and collisions included: their streams are keyed by slot / cell, and the slot order is now a
    function of the input).  Without it two runs agree only statistically."""
    w = wl.c5_channel(12, 8, 8, ppc=30, jitter=0.1, dt_frac=0.5)
    outs = []
    for rep in range(3):
        g = capi.load_gpu(0)
        try:
            g.set_deterministic(rep < 2)
            wl.apply(g, w)
            nc, _ = g.run_steps(1, 6)
            g.sort()
            outs.append((nc, g.download_particles()))
        finally:
            g.close()
    (n0, a), (n1, b), (n2, c) = outs
    assert n0 == n1 and n0 > 1000
    for f in ("X", "Y", "Z", "VX", "VY", "VZ", "IC", "S_ID"):
        assert np.array_equal(a[f], b[f]), f"deterministic runs differ in {f}"
    # the default (atomic-cursor) run is a different realisation of the same physics
    assert len(c) == len(a) and abs(n2 - n0) < 6 * np.sqrt(n0)
    ke = lambda p: float(np.sum(p["VX"] ** 2 + p["VY"] ** 2 + p["VZ"] ** 2))
    assert abs(ke(c) - ke(a)) < 0.02 * ke(a)
