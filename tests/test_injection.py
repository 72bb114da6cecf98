"""Particle creation on the device: INITIAL_SEED, VOLUME_INJECT, BOUNDARIES_EMIT.

Device vs oracle on the same Philox streams (kinds 5, 6, 7: one stream per loop body of the reference), compared as
SETS of particles (the creation order is the order of the append cursor): cells, species and positions bit-identical,
velocities / energies within 1e-12 (libm).  The oracle on the reference's MT19937-64 stream is checked against the
physics: expected counts, Maxwellian moments, the inflow flux of INITINJECTION.
"""
import numpy as np
import pytest

from pantera_pic_dsmc_b200 import capi, workloads as wl

import util

KB = 1.380649e-23


def canon(p):
    """order-independent view: sort by (IC, S_ID, X, Y, Z)"""
    return p[np.lexsort((p["Z"], p["Y"], p["X"], p["S_ID"], p["IC"]))]


def assert_same_set(a, b, what, exact_positions=True):
    if not exact_positions:   # after a flight the positions carry the 1e-12 of the velocities: order by rounded keys
        key = lambda p: p[np.lexsort((np.round(p["Y"], 9), np.round(p["X"], 9), p["S_ID"], p["IC"]))]
        a, b = key(a), key(b)
        assert len(a) == len(b) and np.array_equal(a["IC"], b["IC"]) and np.array_equal(a["S_ID"], b["S_ID"]), what
        for f in ("X", "Y", "Z", "VX", "VY", "VZ"):
            assert np.allclose(a[f], b[f], rtol=1e-10, atol=1e-12), (what, f)
        return
    a, b = canon(a), canon(b)
    assert len(a) == len(b) and len(a) > 0, (what, len(a), len(b))
    for f in ("IC", "S_ID", "X", "Y", "Z"):
        assert np.array_equal(a[f], b[f]), (what, f)
    for f in ("VX", "VY", "VZ"):
        # (a thermal component is ro*sin(teta): next to a zero of the sine the libm difference is absolute, not relative)
        assert np.allclose(a[f], b[f], rtol=1e-12, atol=1e-10), (what, f)
    for f in ("EROT", "EVIB", "DTRIM"):
        assert np.allclose(a[f], b[f], rtol=1e-12, atol=1e-300), (what, f)


def tet_case():
    w = wl.c5_channel(5, 4, 3, ppc=10, jitter=0.1)
    sp = wl.load_species("air", ["N2", "O2"])
    w.species = sp
    w.particles = w.particles[:0]
    w.cfg.collision_type = capi.NO_COLL
    w.cfg.capacity = 200_000
    w.cfg.track_id = 1
    w.bcs["particle_bc"] = [capi.VACUUM, capi.VACUUM, capi.SPECULAR, capi.SPECULAR, capi.SPECULAR, capi.SPECULAR]
    vol = float(np.sum(w.mesh.cell_volumes))
    w.cfg.fnum = 1e20 * vol / 40_000
    return w


def tri_case(axi):
    w = wl.c3_tri_box(nx=6, ny=5, n=10, axi=axi, lo=(0.0, 0.01 if axi else 0.0), hi=(0.1, 0.06), dt=2e-6,
                      wall_bc=("vacuum", "vacuum", "specular", "specular"))
    w.particles = w.particles[:0]
    w.cfg.capacity = 200_000
    w.cfg.track_id = 1
    vol = float(np.sum(w.mesh.cell_volumes))
    w.cfg.fnum = 1e20 * vol / 30_000
    return w


MIX = [[(1, 0.79), (2, 0.21)], [(2, 1.0)]]


def seed_tasks():
    return [dict(mix_id=0, rate=1e20, ux=120.0, uy=-30.0, uz=10.0, ttrax=300.0, ttray=450.0, ttraz=600.0, trot=350.0, tvib=0.0),
            dict(mix_id=1, rate=2e19, ux=0.0, uy=0.0, uz=0.0, ttrax=900.0, ttray=900.0, ttraz=900.0, trot=900.0, tvib=1200.0)]


def configure(lib, w, nranks=1):
    wl.apply(lib, w)
    lib.set_mixtures(MIX)
    lib.set_species_weights([1.0, 2.0])


def run_sources(lib, w, with_emit=True):
    configure(lib, w)
    lib.set_initial_particles_tasks(seed_tasks())
    lib.initial_seed()
    n_seed = lib.num_particles()
    lib.set_volume_inject_tasks([dict(mix_id=0, rate=3e24, ux=0.0, uy=50.0, uz=0.0, ttrax=500.0, ttray=500.0, ttraz=500.0, trot=500.0, tvib=500.0)])
    lib.volume_inject(3)
    n_vinj = lib.num_particles() - n_seed
    n_emit = 0
    if with_emit:
        tasks = wl.emit_tasks_for_group(w.mesh, w.cfg, w.species, MIX[0], pg=1, nrho=2e20, u=(400.0, 20.0, 0.0), ttra=350.0, trot=350.0, tvib=350.0)
        tasks += wl.emit_tasks_for_group(w.mesh, w.cfg, w.species, MIX[1], pg=2, nrho=1e20, u=(-100.0, 0.0, 0.0), ttra=0.0 if w.mesh.dims == 3 else 600.0,
                                         trot=300.0, tvib=300.0, mix_id=1)
        lib.set_emit_tasks(tasks)
        lib.boundaries_emit(4)
        n_emit = lib.num_particles() - n_seed - n_vinj
    return lib.download_particles(), (n_seed, n_vinj, n_emit)


def test_oracle_sources_have_the_expected_counts_and_moments(oracle):
    """MT stream (the reference's): counts within Poisson-like bounds of nrho*V/Fnum, velocity moments of the Maxwellian,
    emitted particles inside their face with inward velocity and DTRIM in [0, DT]."""
    w = tet_case()
    p, (n_seed, n_vinj, n_emit) = run_sources(oracle, w)
    vol = float(np.sum(w.mesh.cell_volumes))
    exp_seed = 1e20 * vol / w.cfg.fnum * (0.79 / 1.0 + 0.21 / 2.0) + 2e19 * vol / w.cfg.fnum / 2.0
    assert abs(n_seed - exp_seed) < 5 * np.sqrt(exp_seed)
    exp_vinj = w.cfg.dt * 3e24 * vol / w.cfg.fnum * (0.79 + 0.21 / 2.0)
    assert abs(n_vinj - exp_vinj) < 5 * np.sqrt(exp_vinj) + 0.5 * w.mesh.nc
    s = p[:n_seed]
    n2 = s[(s["S_ID"] == 1)]
    m = w.species.mass[0]
    assert n2["VX"].mean() == pytest.approx(120.0, abs=5 * np.sqrt(KB * 300 / m / len(n2)) + 3.0)
    for f, T in (("VX", 300.0), ("VY", 450.0), ("VZ", 600.0)):
        assert m * n2[f].var() / KB == pytest.approx(T, rel=0.06)
    assert n2["EROT"].mean() / KB == pytest.approx(350.0, rel=0.05) and not n2["EVIB"].any()
    # containment: every seeded particle lies in its tetrahedron
    mesh = w.mesh
    fn, cn, xyz = np.asarray(mesh.face_normal), np.asarray(mesh.cell_nodes), np.asarray(mesh.node_coords)
    q = p[: n_seed + n_vinj]
    for j in range(4):
        P = xyz[:, cn[j, q["IC"] - 1] - 1]
        dist = (P[0] - q["X"]) * fn[0, j, q["IC"] - 1] + (P[1] - q["Y"]) * fn[1, j, q["IC"] - 1] + (P[2] - q["Z"]) * fn[2, j, q["IC"] - 1]
        assert dist.min() > -1e-12
    e = p[n_seed + n_vinj:]
    assert n_emit > 100 and np.all((e["DTRIM"] >= 0) & (e["DTRIM"] <= w.cfg.dt))
    inflow = e[e["S_ID"] == 1]
    assert np.all(np.abs(inflow["X"]) < 1e-12) and np.all(inflow["VX"] > 0)      # group 1 = xmin face, emitted inwards
    exp_emit = sum(sum(t["nfs"]) for t in wl.emit_tasks_for_group(w.mesh, w.cfg, w.species, MIX[0], 1, 2e20, (400.0, 20.0, 0.0), 350.0))
    n_in = int(np.sum(np.abs(e["X"]) < 1e-12))
    assert abs(n_in - exp_emit) < 5 * np.sqrt(exp_emit) + 0.5 * 24


@pytest.mark.gpu
@pytest.mark.parametrize("case", ["tets", "tris", "axi"])
def test_device_sources_match_oracle_value_for_value(gpu, oracle_philox, case):
    w = tet_case() if case == "tets" else tri_case(case == "axi")
    a, ca = run_sources(oracle_philox, w)
    b, cb = run_sources(gpu, w)
    assert ca == cb and min(ca) > 50, (ca, cb)
    assert_same_set(a, b, case)
    # and the created particles fly: emitted ones use their own DTRIM in the first ADVECT
    gpu.advect(5)
    oracle_philox.advect(5)
    pa, pb = oracle_philox.download_particles(), gpu.download_particles()
    assert len(pa) == len(pb)
    assert_same_set(pa, pb, case + " after advect", exact_positions=False)


@pytest.mark.gpu
def test_device_structured_seed_and_inject_match_oracle(gpu, oracle_philox):
    w = wl.c1_argon_box(10, nx=9, ny=7)
    w.species = wl.load_species("air", ["N2", "O2"])
    w.particles = w.particles[:0]
    w.cfg.capacity = 200_000
    w.cfg.fnum = 1e20 / 50_000
    for lib in (oracle_philox, gpu):
        configure(lib, w)
        lib.set_initial_particles_tasks(seed_tasks())
        lib.initial_seed()
        lib.set_volume_inject_tasks([dict(mix_id=1, rate=4e25, ux=0.0, uy=0.0, uz=0.0, ttrax=300.0, ttray=300.0, ttraz=300.0, trot=300.0, tvib=300.0)])
        lib.volume_inject(9)
    a, b = oracle_philox.download_particles(), gpu.download_particles()
    assert len(a) > 40_000
    assert_same_set(a, b, "structured")


@pytest.mark.gpu
def test_open_channel_runs_with_device_injection(gpu):
    """C5 open-channel variant: vacuum end caps, Boundary_emit at x = 0, every step on the device:
    BOUNDARIES_EMIT -> ADVECT -> sort -> DSMC.  The population settles where inflow balances outflow."""
    w = wl.c5_channel(10, 4, 4, ppc=20, wall_bc=("vacuum", "vacuum", "specular", "specular", "specular", "specular"), u_drift=300.0)
    w.cfg.capacity = 4 * len(w.particles)
    wl.apply(gpu, w)
    gpu.set_mixtures([[(1, 1.0)]])
    tasks = wl.emit_tasks_for_group(w.mesh, w.cfg, w.species, [(1, 1.0)], pg=1, nrho=1e20, u=(300.0, 0.0, 0.0), ttra=300.0)
    gpu.set_emit_tasks(tasks)
    n0 = gpu.num_particles()
    counts = []
    for t in range(1, 41):
        gpu.boundaries_emit(t)
        gpu.run_steps(t, 1)
        counts.append(gpu.num_particles())
    assert 0.5 * n0 < counts[-1] < 1.3 * n0
    assert abs(counts[-1] - counts[-10]) < 0.1 * n0
    p = gpu.download_particles()
    assert p["VX"].mean() > 150.0
