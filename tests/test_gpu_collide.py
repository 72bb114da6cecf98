"""Parity of the CUDA sort + DSMC collisions + MCC + deposit with the CPU oracle, through the C ABI.

Collisions: the reference's own RNG stream (one sequential MT19937-64) cannot be reproduced by a parallel
code, so north_star asks for statistical agreement.  We do better where it is cheap: the oracle can run on
the device's counter-based Philox streams (rng_mode 1), and the device keeps the reference's sequential
pair semantics inside a cell exactly, so on the same sorted input the two agree value for value
(rtol 1e-12; libm differs by <= 2 ulp).  Statistical checks against kinetic theory cover the rest.
"""
import numpy as np
import pytest

from pantera_pic_dsmc_b200 import capi, datafiles as df, workloads as wl

import util

pytestmark = pytest.mark.gpu
KB = df.KB


def sorted_pair(gpu, orc, w):
    """Sort on the device, hand the sorted particles to the oracle: both then index cells identically."""
    w.cfg.track_id = 1
    wl.apply(gpu, w)
    gpu.sort()
    ps = gpu.download_particles()
    assert np.all(np.diff(ps["IC"]) >= 0), "store not sorted by cell"
    wl.apply(orc, w, upload=False)
    orc.upload_particles(ps)
    return ps


def test_sort_is_a_permutation_grouped_by_cell(gpu):
    w = wl.c1_argon_box(200_000, nx=32, ny=16)
    w.cfg.track_id = 1
    wl.apply(gpu, w)
    gpu.sort()
    ps = gpu.download_particles()
    assert np.all(np.diff(ps["IC"]) >= 0)
    a, b = util.by_id(w.particles), util.by_id(ps)
    util.assert_particles_match(a, b, exact_fields=util.FLOAT_FIELDS, what="sort permutation")
    npc = gpu.get_cell_counts(32 * 16)
    assert np.array_equal(npc, np.bincount(w.particles["IC"] - 1, minlength=512))


def test_sort_compacts_removed_particles(gpu, oracle_philox):
    w = wl.c3_tri_box(n=40_000, wall_bc=("vacuum",) * 4, u_drift=(1500.0, 0, 0), dt=1e-5)
    util.setup_pair(gpu, oracle_philox, w)
    gpu.advect(1)
    oracle_philox.advect(1)
    n = gpu.num_particles()
    assert n == oracle_philox.num_particles() < 40_000
    gpu.sort()
    ps = gpu.download_particles()
    assert len(ps) == n and np.all(ps["IC"] >= 1) and np.all(np.diff(ps["IC"]) >= 0)


@pytest.mark.parametrize("nx,ny,n", [(1, 1, 3000), (1, 1, 60_000), (16, 8, 100_000), (1, 2, 5)])
def test_vss_collisions_value_for_value(gpu, oracle_philox, nx, ny, n):
    w = wl.c1_argon_box(n, nx=nx, ny=ny)
    sorted_pair(gpu, oracle_philox, w)
    for tID in (1, 2, 3):
        co = oracle_philox.collide(tID)
        cg = gpu.collide(tID)
        assert co == cg, f"TIMESTEP_COLL/REAC differ at step {tID}: oracle {co} device {cg}"
    a, b = oracle_philox.download_particles(), gpu.download_particles()
    util.assert_particles_match(a, b, rtol=1e-12, atol=1e-9, what="VSS argon")


@pytest.mark.parametrize("nx,ny,n", [(16, 8, 100_000), (2, 1, 40_000)])
def test_deferred_velocity_reorder_value_for_value(oracle_philox, monkeypatch, nx, ny, n):
    """PANTERA_GPU_DEFER_VEL=1: the sort moves pos only, the collision kernels (CELL path, and the GLOBAL path for the
    20000-particle cells of the second case) fetch the velocities through the source index.  Same results, and a
    download between a sort and a collide resolves the pending reorder."""
    monkeypatch.setenv("PANTERA_GPU_DEFER_VEL", "1")
    gpu = capi.load_gpu(0)
    try:
        w = wl.c1_argon_box(n, nx=nx, ny=ny)
        w.cfg.track_id = 1
        wl.apply(gpu, w)
        wl.apply(oracle_philox, w, upload=False)
        for tID in (1, 2):
            gpu.advect(tID)
            gpu.sort()
            ps = gpu.download_particles()
            gpu.advect(10 + tID)                         # leaves the store unsorted: pgpu_collide sorts (deferred) itself
            ps = gpu.download_particles()
            gpu.upload_particles(ps)
            cg = gpu.collide(tID)                        # sort with deferred vel + gather in the collision kernels
            mid = gpu.download_particles()
            order = np.argsort(mid["IC"], kind="stable")
            assert np.all(np.diff(mid["IC"]) >= 0)
            # the oracle collides the device's sorted order, from the pre-collision velocities
            pre = util.by_id(ps)
            srt = pre[np.searchsorted(pre["ID"], mid["ID"])]
            srt["IC"] = mid["IC"]
            oracle_philox.upload_particles(srt)
            co = oracle_philox.collide(tID)
            assert co == cg and cg[0] > 0, (co, cg)
            util.assert_particles_match(oracle_philox.download_particles(), mid, rtol=1e-12, atol=1e-9, what="deferred vel")
    finally:
        gpu.close()


def test_air_bl_tce_value_for_value(gpu, oracle_philox):
    w = wl.c2_air_box(60_000, nx=6, ny=5, T=12000.0, dt=3e-7)
    sorted_pair(gpu, oracle_philox, w)
    n0 = gpu.num_particles()
    tot_o, tot_g = np.zeros(2, dtype=np.int64), np.zeros(2, dtype=np.int64)
    for tID in (1, 2):
        tot_o += oracle_philox.collide(tID)
        tot_g += gpu.collide(tID)
        # products are appended at the end in both codes, but in different order: realign through a sort
        if tID == 1:
            po, pg = oracle_philox.download_particles(), gpu.download_particles()
            assert len(po) == len(pg) > n0
            util.assert_particles_match(po[:n0], pg[:n0], rtol=1e-12, atol=1e-9, what="air step 1 (existing)")
            key = lambda p: np.lexsort((p["VZ"], p["VY"], p["VX"], p["X"]))
            ao, ag = po[n0:], pg[n0:]
            ao, ag = ao[key(ao)], ag[key(ag)]
            for f in ("X", "Y", "Z", "S_ID", "IC"):
                assert np.array_equal(ao[f], ag[f])
            for f in ("VX", "VY", "VZ", "EROT", "EVIB"):
                assert np.allclose(ao[f], ag[f], rtol=1e-11, atol=1e-9 if f[0] == "V" else 1e-32)
            break
    assert np.array_equal(tot_o, tot_g) and tot_g[1] > 0


def test_collisions_after_moves_conserve_and_match_rate(gpu):
    """Full device loop (ADVECT -> sort -> collide) on the tet channel: energy, momentum and the
    kinetic-theory collision rate with specular walls."""
    w = wl.c5_channel(10, 6, 6, ppc=40, wall_bc=("specular",) * 6, u_drift=0.0, T0=400.0)
    wl.apply(gpu, w)
    m = w.species.mass
    e0 = util.kinetic_energy(w.particles, m)
    n, steps = len(w.particles), 10
    ncoll, nreac = gpu.run_steps(1, steps)
    q = gpu.download_particles()
    assert len(q) == n and nreac == 0
    assert util.kinetic_energy(q, m) == pytest.approx(e0, rel=1e-11)
    expect = 0.5 * 0.2 * n * steps   # N nu dt / 2 with dt = 0.2/nu
    assert abs(ncoll - expect) < 0.04 * expect, (ncoll, expect)


def test_heat_bath_relaxation_matches_oracle_statistically(gpu, oracle):
    """C1: anisotropic Maxwellian (600,300,300) K -> 400 K.  Device (Philox) vs oracle on the REFERENCE stream
    (MT19937-64): temperatures agree within 3-sigma sampling error at every sampled step."""
    n = 40_000
    w = wl.c1_argon_box(n, nx=2, ny=2)
    wl.apply(gpu, w)
    wl.apply(oracle, w)
    m = w.species.mass[0]
    sig = 400.0 * np.sqrt(2.0 / n)
    for k in range(4):
        gpu.run_steps(1 + 5 * k, 5)
        oracle.run_steps(1 + 5 * k, 5)
        tg = np.array([m * gpu.download_particles()[f].var() / KB for f in ("VX", "VY", "VZ")])
        to = np.array([m * oracle.download_particles()[f].var() / KB for f in ("VX", "VY", "VZ")])
        assert np.all(np.abs(tg - to) < 3 * np.sqrt(2) * sig * 1.5), (k, tg, to)
    assert abs(tg[0] - tg[1]) < 0.35 * (600 - 300)


def test_mcc_vahedi_value_for_value(gpu, oracle_philox):
    w = wl.c4_ccp_1d(32, 40_000)
    w.cfg.dt = 2e-9      # P_null = 1 - exp(-2) so that most particles are tested
    w.particles["DTRIM"] = w.cfg.dt
    util.setup_pair(gpu, oracle_philox, w)
    n0 = len(w.particles)
    co, cg = oracle_philox.mcc_vahedi(5), gpu.mcc_vahedi(5)
    assert co == cg and cg[1] > 100
    po, pg = oracle_philox.download_particles(), gpu.download_particles()
    assert len(po) == len(pg) > n0
    # existing particles keep their slots (download sorts by cell on the device only if tombstones exist)
    util.assert_particles_match(util.by_id(po[:n0]), util.by_id(pg[pg["ID"] < (n0 << 8)]), rtol=1e-11, atol=1e-9, what="MCC existing")
    key = lambda p: np.lexsort((p["VZ"], p["VY"], p["VX"], p["X"]))
    ao, ag = po[n0:], pg[pg["ID"] >= (n0 << 8)]
    ao, ag = ao[key(ao)], ag[key(ag)]
    assert np.array_equal(ao["S_ID"], ag["S_ID"]) and np.array_equal(ao["X"], ag["X"])
    for f in ("VX", "VY", "VZ"):
        assert np.allclose(ao[f], ag[f], rtol=1e-10, atol=1e-7)


@pytest.mark.parametrize("kind,cell_density", [("ar", False), ("ar", True), ("air", False)])
def test_mcc_collisions_value_for_value(gpu, oracle_philox, kind, cell_density):
    """MCC_COLLISIONS (collisions.f90:1261-1595): same Philox stream per particle on both
    sides -> the same particles collide, the same partners/products are created, velocities and internal
    energies agree to 1e-10 (libm pow/exp/sincos differ by a few ulp)."""
    if kind == "ar":
        w = wl.mcc_box("ar", n=40_000, cell_density=cell_density)
    else:
        w = wl.mcc_box("air", n=40_000, T=15000.0, bg_dens=5e21)
    w.cfg.capacity = 5 * len(w.particles)   # room for one partner per background component + third products
    util.setup_pair(gpu, oracle_philox, w)
    n0 = len(w.particles)
    co, cg = oracle_philox.mcc_collisions(7), gpu.mcc_collisions(7)
    assert co == cg and cg[0] > 1000
    if kind == "air":
        assert cg[1] > 20
    po, pg = oracle_philox.download_particles(), gpu.download_particles()
    assert len(po) == len(pg) > n0
    util.assert_particles_match(util.by_id(po[:n0]), util.by_id(pg[pg["ID"] < (n0 << 8)]), rtol=1e-10, atol=1e-9,
                                what="MCC-VSS existing")
    key = lambda p: np.lexsort((p["VZ"], p["VY"], p["VX"], p["S_ID"], p["X"]))
    ao, ag = po[n0:], pg[pg["ID"] >= (n0 << 8)]
    ao, ag = ao[key(ao)], ag[key(ag)]
    assert np.array_equal(ao["S_ID"], ag["S_ID"]) and np.array_equal(ao["X"], ag["X"]) and np.array_equal(ao["IC"], ag["IC"])
    for f in ("VX", "VY", "VZ", "EROT", "EVIB"):
        assert np.allclose(ao[f], ag[f], rtol=1e-9, atol=1e-7 if f[0] == "V" else 1e-30), f


def test_run_steps_mcc_collision_type(gpu, oracle_philox):
    """Collision_type MCC inside the step loop (timecycle.f90:288-296): ADVECT -> sort -> MCC_COLLISIONS."""
    w = wl.mcc_box("ar", n=20_000)
    util.setup_pair(gpu, oracle_philox, w)
    cg = gpu.run_steps(1, 2)
    assert cg[0] > 100 and cg[1] == 0
    assert gpu.num_particles() == len(w.particles) + cg[0]


def test_remove_species(gpu, oracle_philox):
    w = wl.c4_ccp_1d(16, 5000)
    util.setup_pair(gpu, oracle_philox, w)
    sp = w.species.id("e")
    gpu.remove_species(sp)
    oracle_philox.remove_species(sp)
    a, b = util.by_id(oracle_philox.download_particles()), util.by_id(gpu.download_particles())
    util.assert_particles_match(a, b, exact_fields=util.FLOAT_FIELDS, what="remove species")
    assert not np.any(b["S_ID"] == sp)


@pytest.mark.parametrize("kind", ["line", "tri", "tet", "cic"])
def test_deposit_charge_within_1e12(gpu, oracle, kind):
    """Charge: |device - compensated oracle sum| <= 1e-12 * sum|terms| per node."""
    if kind == "line":
        w = wl.c4_ccp_1d(64, 100_000)
    elif kind == "tri":
        w = wl.c3_tri_box(nx=12, ny=9, n=80_000, species=("N2", "O2"))
        w.species.charge[:] = [1.0, -1.0]
    elif kind == "tet":
        w = wl.c5_channel(5, 4, 3, ppc=60)
        w.species.charge[:] = 1.0
    else:
        sp = wl.load_species("ar", ["Ar+", "e"])
        n, nx, ny = 100_000, 16, 12
        cfg = capi.make_config(dims=2, nx=nx, ny=ny, fnum=1e8)
        rng = np.random.default_rng(7)
        p = wl.new_particles(n)
        p["X"], p["Y"] = rng.random(n), rng.random(n)
        p["S_ID"] = rng.integers(1, 3, n)
        p["IC"] = (p["X"] * nx).astype(np.int32) + nx * (p["Y"] * ny).astype(np.int32) + 1
        p["DTRIM"] = cfg.dt
        w = wl.Workload("cic", cfg, sp, p)
    wl.apply(gpu, w)
    wl.apply(oracle, w)
    nn = w.mesh.nn if w.mesh is not None else (w.cfg.nx + 1) * (w.cfg.ny + 1)
    if w.mesh is not None:
        sc = np.linspace(-1.0, 1.0, nn) * 1e3
        gpu.set_surface_charge(sc)
        oracle.set_surface_charge(sc)
    ref, ab = oracle.deposit_charge_compensated(nn)
    got = gpu.deposit_charge(nn)
    assert np.abs(ref).max() > 0
    assert np.all(np.abs(got - ref) <= 1e-12 * ab + 1e-300), np.max(np.abs(got - ref) / (ab + 1e-300))


def test_borgnakke_larsen_equipartition_on_device(gpu):
    """N2/O2 heat bath started with cold internal modes: T_rot -> T_tr, total energy conserved (the oracle's version of this
    test runs on the reference's MT stream: tests/test_oracle_physics.py)."""
    w = wl.c2_air_box(40_000, nx=4, ny=4, T=2000.0, with_reactions=False, species=("O2", "N2"), frac=(0.5, 0.5), dt=4e-7)
    w.particles["EROT"] = 0.0
    w.particles["EVIB"] = 0.0
    wl.apply(gpu, w)
    m = w.species.mass
    etot0 = util.kinetic_energy(w.particles, m)
    gpu.run_steps(1, 150)
    q = gpu.download_particles()
    etot1 = util.kinetic_energy(q, m) + q["EROT"].sum() + q["EVIB"].sum()
    assert etot1 == pytest.approx(etot0, rel=1e-10)
    ttr = 2.0 / 3.0 * util.kinetic_energy(q, m) / len(q) / KB
    trot = q["EROT"].mean() / KB
    assert abs(trot - ttr) < 0.04 * ttr


@pytest.mark.parametrize("T", [8000.0, 12000.0])
def test_tce_recovers_the_arrhenius_rate_on_device(gpu, T):
    """Dissociations of N2 + O2 -> N2 + O + O in an equilibrium gas against k(T) = A T^n exp(-Ea/kT) (data/air.tce)."""
    from test_oracle_physics import arrhenius_case
    w, per_step = arrhenius_case(T)
    wl.apply(gpu, w)
    steps = 6
    _, nreac = gpu.run_steps(1, steps)
    ratio = nreac / (per_step * steps)
    assert 0.85 < ratio < 1.0 + 4.0 / np.sqrt(per_step * steps) + 0.03, (T, nreac, per_step * steps, ratio)
