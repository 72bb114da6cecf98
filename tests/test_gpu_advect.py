"""Parity of the CUDA ADVECT (pgpu_advect) with the CPU oracle, through the C ABI.

Bar (BASELINE.json north_star): cell indices bit-exact, positions within 1e-12 relative in FP64.  Away from
stochastic walls the device arithmetic is the same IEEE sequence as the oracle's, so those cases are required
to be bit-identical; wall re-emission goes through libm (log/sin/cos/exp differ by <= 2 ulp between glibc and
CUDA), so those particles are compared at rtol 1e-12 on the SAME Philox streams (oracle rng_mode 1).
"""
import numpy as np
import pytest

from pantera_pic_dsmc_b200 import capi, meshgen as mg, workloads as wl

import util

pytestmark = pytest.mark.gpu

MOVE_FIELDS = ("X", "Y", "Z", "VX", "VY", "VZ", "EROT", "EVIB")


def run_advect_pair(gpu, orc, w, steps=1, tID0=1, resync=False):
    """resync: the wall-RNG stream of a particle is keyed by its slot.  When particles are removed the oracle
    compacts by swap-with-last (REMOVE_PARTICLE_ARRAY, particle.f90:142-179) while the device leaves tombstones
    until the next sort, so the slot numbering diverges; both stores are then re-seeded with the (already
    compared) device particles before the next step."""
    util.setup_pair(gpu, orc, w)
    for s in range(steps):
        orc.advect(tID0 + s)
        gpu.advect(tID0 + s)
        if resync and s + 1 < steps:
            a, b = util.by_id(orc.download_particles()), util.by_id(gpu.download_particles())
            assert np.array_equal(a["ID"], b["ID"]), f"different sets of surviving particles after step {s}"
            util.assert_particles_match(a, b, rtol=1e-12, what=f"resync step {s}")
            orc.upload_particles(b)
            gpu.upload_particles(b)
    a, b = util.by_id(orc.download_particles()), util.by_id(gpu.download_particles())
    assert np.array_equal(a["ID"], b["ID"]), "different sets of surviving particles"
    return a, b


@pytest.mark.parametrize("nx,ny", [(1, 1), (7, 5), (64, 64)])
def test_structured_periodic_bit_exact(gpu, oracle_philox, nx, ny):
    w = wl.c1_argon_box(50_000, nx=nx, ny=ny)
    w.cfg.collision_type = capi.NO_COLL
    w.cfg.dt = 3e-4  # several wraps per step for the fast particles
    w.particles["DTRIM"] = w.cfg.dt
    a, b = run_advect_pair(gpu, oracle_philox, w, steps=3)
    util.assert_particles_match(a, b, exact_fields=MOVE_FIELDS, what="periodic box")
    bo, _ = oracle_philox.get_boundary_counts(1)
    bg, _ = gpu.get_boundary_counts(1)
    assert np.array_equal(bo, bg) and bo.sum() > 0


def test_structured_specular_bit_exact(gpu, oracle_philox):
    w = wl.c1_argon_box(30_000, nx=9, ny=4, periodic=False)
    w.cfg.collision_type = capi.NO_COLL
    w.cfg.dt = 5e-4
    w.particles["DTRIM"] = w.cfg.dt
    a, b = run_advect_pair(gpu, oracle_philox, w, steps=2)
    util.assert_particles_match(a, b, exact_fields=MOVE_FIELDS, what="specular box")


def test_structured_particle_dtrim_is_honoured(gpu, oracle_philox):
    """Freshly injected particles carry DTRIM = rf()*DT (timecycle.f90:463): first step uses it."""
    w = wl.c1_argon_box(5_000, nx=4, ny=4)
    w.cfg.collision_type = capi.NO_COLL
    rng = np.random.default_rng(0)
    w.particles["DTRIM"] = rng.random(len(w.particles)) * w.cfg.dt
    a, b = run_advect_pair(gpu, oracle_philox, w, steps=2)
    util.assert_particles_match(a, b, exact_fields=MOVE_FIELDS, what="per-particle DTRIM")


def test_structured_diffuse_react_walls(gpu, oracle_philox):
    sp_names = ("N2", "O2")
    w = wl.c2_air_box(20_000, nx=5, ny=3, T=3000.0, species=sp_names, frac=(0.5, 0.5), with_reactions=False)
    w.cfg.collision_type = capi.NO_COLL
    w.cfg.dt = 2e-4
    w.particles["DTRIM"] = w.cfg.dt
    for i, v in enumerate([0, 0, 0, 0]):
        w.cfg.bool_periodic[i] = v
    for i, v in enumerate([1, 0, 1, 0]):
        w.cfg.bool_diffuse[i] = v
    for i, v in enumerate([0, 1, 0, 0]):
        w.cfg.bool_specular[i] = v       # side 4 (ymax) has no BC: particles are removed there
    for i, v in enumerate([1, 1, 0, 0]):
        w.cfg.bool_react[i] = v
    w.cfg.boundtemp = 800.0
    w.wall_reactions = dict(r_sp=[1, 1, 2], p1_sp=[2, -1, -1], nprod=[1, 0, 0], prob=[0.3, 0.2, 0.1])
    a, b = run_advect_pair(gpu, oracle_philox, w, steps=2, resync=True)
    assert len(a) < 20_000
    util.assert_particles_match(a, b, rtol=1e-12, what="diffuse/react box")
    assert util.exact_fraction(a, b) > 0.5
    bo, _ = oracle_philox.get_boundary_counts(2)
    bg, _ = gpu.get_boundary_counts(2)
    assert np.array_equal(bo, bg)


def test_structured_interior_walls(gpu, oracle_philox):
    w = wl.c1_argon_box(20_000, nx=6, ny=6)
    w.cfg.collision_type = capi.NO_COLL
    w.cfg.dt = 2e-4
    w.particles["DTRIM"] = w.cfg.dt
    # three segments: specular, diffuse, porous-specular.  Normal = (y2-y1, x1-x2)/len convention of DEF_WALL
    x1, y1, x2, y2 = [0.3, 0.6, 0.2], [0.2, 0.8, 0.5], [0.3, 0.9, 0.5], [0.7, 0.8, 0.5]
    nx_, ny_ = [], []
    for a_, b_, c_, d_ in zip(x1, y1, x2, y2):
        ln = np.hypot(c_ - a_, d_ - b_)
        nx_.append((d_ - b_) / ln)
        ny_.append((a_ - c_) / ln)
    w.walls = dict(x1=x1, y1=y1, x2=x2, y2=y2, normx=nx_, normy=ny_, temp=[300.0, 600.0, 300.0], specular=[1, 0, 1],
                   diffuse=[0, 1, 0], porous=[0, 0, 1], transmissivity=[0.0, 0.0, 0.4], react=[0, 0, 0])
    a, b = run_advect_pair(gpu, oracle_philox, w, steps=2, resync=True)
    util.assert_particles_match(a, b, rtol=1e-12, what="interior walls")
    _, wo = oracle_philox.get_boundary_counts(1, 3)
    _, wg = gpu.get_boundary_counts(1, 3)
    assert np.array_equal(wo, wg) and wo.sum() > 0


def test_nonuniform_grid_reindex(gpu, oracle_philox):
    w = wl.c1_argon_box(20_000, nx=6, ny=5)
    w.cfg.collision_type = capi.NO_COLL
    w.cfg.grid_type = capi.RECTILINEAR_NONUNIFORM
    xc = np.array([0.0, 0.05, 0.15, 0.3, 0.5, 0.75, 1.0])
    yc = np.array([0.0, 0.1, 0.3, 0.6, 0.8, 1.0])
    util.setup_pair(gpu, oracle_philox, w)
    for lib in (gpu, oracle_philox):
        lib.set_grid_coords(xc, yc)
        lib.advect(1)
    a, b = util.by_id(oracle_philox.download_particles()), util.by_id(gpu.download_particles())
    util.assert_particles_match(a, b, exact_fields=MOVE_FIELDS, what="non-uniform grid")
    ix = np.searchsorted(xc, a["X"], side="right")
    iy = np.searchsorted(yc, a["Y"], side="right")
    assert np.array_equal(a["IC"], ix + 6 * (iy - 1))


@pytest.mark.parametrize("jitter", [0.0, 0.25])
def test_tets_specular_bit_exact(gpu, oracle_philox, jitter):
    w = wl.c5_channel(8, 5, 4, ppc=30, jitter=jitter, wall_bc=("specular",) * 6, dt_frac=0.6)
    w.cfg.collision_type = capi.NO_COLL
    a, b = run_advect_pair(gpu, oracle_philox, w, steps=3)
    util.assert_particles_match(a, b, exact_fields=MOVE_FIELDS, what="tets specular")
    assert np.any(a["IC"] != util.by_id(w.particles)["IC"])


def test_tets_all_wall_models_and_tallies(gpu, oracle_philox):
    w = wl.c5_channel(6, 4, 4, ppc=40, jitter=0.15, dt_frac=0.8,
                      wall_bc=("vacuum", "specular", "diffuse", "cll", "diffuse", "specular"), wall_temp=500.0)
    w.cfg.collision_type = capi.NO_COLL
    w.cfg.dump_bound_avg_every = 1
    a, b = run_advect_pair(gpu, oracle_philox, w, steps=2, resync=True)
    assert len(a) < len(w.particles)
    util.assert_particles_match(a, b, rtol=1e-12, what="tets walls")
    assert util.exact_fraction(a, b) > 0.6
    to = oracle_philox.get_boundary_tallies(w.mesh.nboundcells, 1)
    tg = gpu.get_boundary_tallies(w.mesh.nboundcells, 1)
    assert np.abs(to).sum() > 0
    scale = np.abs(to).max(axis=(1, 2), keepdims=True) + 1e-300
    assert np.all(np.abs(to - tg) <= 1e-11 * scale)


@pytest.mark.parametrize("axi", [False, True])
def test_triangles_walls_reactions(gpu, oracle_philox, axi):
    w = wl.c3_tri_box(nx=9, ny=7, n=30_000, axi=axi, lo=(0.0, 0.01 if axi else 0.0), hi=(0.1, 0.06), dt=4e-6,
                      wall_bc=("vacuum", "specular", "cll", "diffuse"), wall_reactions=True, react=(0, 1, 1, 1))
    a, b = run_advect_pair(gpu, oracle_philox, w, steps=2, tID0=4, resync=True)
    util.assert_particles_match(a, b, rtol=1e-12, atol=1e-18, what=f"triangles axi={axi}")
    assert len(a) < len(w.particles)
    to = oracle_philox.get_boundary_tallies(w.mesh.nboundcells, 2)
    tg = gpu.get_boundary_tallies(w.mesh.nboundcells, 2)
    scale = np.abs(to).max(axis=(1, 2), keepdims=True) + 1e-300
    assert np.all(np.abs(to - tg) <= 1e-11 * scale)


def test_line_mesh_pic_kick_and_dielectric_charging(gpu, oracle_philox):
    w = wl.c4_ccp_1d(48, 15_000, wall_bc=("diffuse", "vacuum"), field_bc=(capi.DIELECTRIC_BC, capi.DIRICHLET_BC))
    w.cfg.collision_type = capi.NO_COLL
    w.cfg.dt = 4e-10
    w.particles["DTRIM"] = w.cfg.dt
    for i, v in enumerate([0.0, 0.0, 0.02]):
        w.cfg.external_b[i] = v
    a, b = run_advect_pair(gpu, oracle_philox, w, steps=2, resync=True)
    util.assert_particles_match(a, b, rtol=1e-12, atol=1e-20, what="1-D PIC")
    sco = oracle_philox.get_surface_charge(w.mesh.nn)
    scg = gpu.get_surface_charge(w.mesh.nn)
    assert np.abs(sco).sum() > 0
    assert np.allclose(sco, scg, rtol=1e-11, atol=1e-11 * np.abs(sco).max())


def test_structured_pic_cic_gather(gpu, oracle_philox):
    sp = wl.load_species("ar", ["Ar+", "e"])
    n, nx, ny = 20_000, 8, 6
    cfg = capi.make_config(dims=2, nx=nx, ny=ny, dt=2e-10, pic_type=capi.PIC_EXPLICIT, bool_periodic=[1] * 4,
                           bool_z_periodic=1, external_b=[0.01, 0.0, 0.03])
    rng = np.random.default_rng(3)
    p = wl.new_particles(n)
    p["X"], p["Y"], p["Z"] = rng.random(n), rng.random(n), rng.random(n) - 0.5
    p["S_ID"] = rng.integers(1, 3, n)
    v = wl.maxwellian(rng, n, sp.mass[p["S_ID"] - 1], 20000.0)
    p["VX"], p["VY"], p["VZ"] = v[:, 0], v[:, 1], v[:, 2]
    p["DTRIM"] = cfg.dt
    p["IC"] = (p["X"] * nx).astype(np.int32) + nx * (p["Y"] * ny).astype(np.int32) + 1
    ef = np.asfortranarray(rng.standard_normal((3, ny + 1, nx + 1)) * 1e4)
    w = wl.Workload("pic-cic", cfg, sp, p, efield=ef)
    a, b = run_advect_pair(gpu, oracle_philox, w, steps=2)
    util.assert_particles_match(a, b, exact_fields=MOVE_FIELDS, what="structured PIC")


def test_triangles_pic_nodal_bfield(gpu, oracle_philox):
    """APPLY_B_FIELD (src/fields.f90:4227-4249): nodal P1 interpolation of B_FIELD(3,1,NNODES) on a 2-D triangle mesh
    through pgpu_set_bfield, plus a uniform external field and a cell-constant E (APPLY_E_FIELD, unstructured branch).
    No stochastic wall is involved (specular box), so the Boris kick and the flight must be bit-identical."""
    w = wl.c3_tri_box(nx=10, ny=8, n=25_000, lo=(0.0, 0.0), hi=(0.1, 0.08), dt=3e-10, species=("N2", "O2"),
                      wall_bc=("specular",) * 4, T0=20000.0, u_drift=(0.0, 0.0, 0.0))
    sp = wl.load_species("ar", ["Ar+", "e"])
    rng = np.random.default_rng(12)
    n = len(w.particles)
    w.species = sp
    w.particles["S_ID"] = rng.integers(1, 3, n)
    v = wl.maxwellian(rng, n, sp.mass[w.particles["S_ID"] - 1], 20000.0)
    w.particles["VX"], w.particles["VY"], w.particles["VZ"] = v[:, 0], v[:, 1], v[:, 2]
    w.particles["EROT"] = 0.0
    w.particles["EVIB"] = 0.0
    w.cfg.pic_type = capi.PIC_EXPLICIT
    for i, b in enumerate([0.0, 0.005, 0.01]):
        w.cfg.external_b[i] = b
    m = w.mesh
    xy = np.asarray(m.node_coords)
    bf = np.zeros((3, m.nn), order="F")
    bf[0] = 0.02 * np.sin(40.0 * xy[0]) + 0.003 * rng.standard_normal(m.nn)
    bf[1] = 0.02 * np.cos(55.0 * xy[1])
    bf[2] = 0.05 * (1.0 + 10.0 * xy[0] - 7.0 * xy[1])
    w.efield = np.asfortranarray(rng.standard_normal((3, m.nc)) * 2e4)
    util.setup_pair(gpu, oracle_philox, w)
    gpu.set_bfield(bf)
    oracle_philox.set_bfield(bf)
    p0 = util.by_id(w.particles)
    for s in range(3):
        gpu.advect(1 + s)
        oracle_philox.advect(1 + s)
    a, b = util.by_id(oracle_philox.download_particles()), util.by_id(gpu.download_particles())
    util.assert_particles_match(a, b, exact_fields=MOVE_FIELDS, what="2-D PIC with nodal B")
    # the nodal field matters: the same run without it ends elsewhere
    gpu.upload_particles(p0)
    gpu.set_bfield(np.zeros((3, m.nn), order="F"))
    for s in range(3):
        gpu.advect(1 + s)
    c = util.by_id(gpu.download_particles())
    assert np.abs(c["VZ"] - b["VZ"]).max() > 1e-3 * np.abs(b["VZ"]).max()


def test_solid_cells_act_as_walls(gpu, oracle_philox):
    w = wl.c3_tri_box(nx=8, ny=8, n=20_000, lo=(0, 0), hi=(0.08, 0.08), dt=6e-6, wall_bc=("specular",) * 4)
    m = w.mesh
    cen = np.asarray(m.node_coords)[:2, np.asarray(m.cell_nodes) - 1].mean(axis=1)
    solid = (np.abs(cen[0] - 0.04) < 0.011) & (np.abs(cen[1] - 0.04) < 0.011)
    m.cell_pg = np.where(solid, 5, -1).astype(np.int32)
    nb = np.asarray(m.cell_neighbors)
    fpg = np.asarray(m.cell_faces_pg).copy()
    touching = (nb > 0) & solid[np.maximum(nb, 1) - 1] & ~solid[None, :]
    fpg[touching] = 6                      # interior faces against the solid body carry a marker
    m.cell_faces_pg = np.asfortranarray(fpg)
    mg._finish_boundary(m)
    w.bcs = dict(particle_bc=[capi.SPECULAR] * 4 + [capi.VACUUM, capi.DIFFUSE], volume_bc=[capi.FLUID] * 4 + [capi.SOLID, capi.FLUID],
                 wall_temp=[300.0] * 5 + [900.0], acc_n=[1.0] * 6, acc_t=[1.0] * 6, react=[0] * 6)
    keep = ~solid[w.particles["IC"] - 1]
    w.particles = w.particles[keep]
    a, b = run_advect_pair(gpu, oracle_philox, w, steps=3)
    util.assert_particles_match(a, b, rtol=1e-12, atol=1e-18, what="solid body")
    assert not np.any(solid[a["IC"] - 1])


def test_empty_store_and_error_paths(gpu):
    w = wl.c1_argon_box(100)
    w.particles = w.particles[:0]
    wl.apply(gpu, w)
    gpu.advect(1)
    assert gpu.num_particles() == 0
    assert len(gpu.download_particles()) == 0
    bad = wl.c1_argon_box(100, nx=2, ny=2)
    bad.particles["IC"][3] = 99
    with pytest.raises(capi.PanteraError):
        wl.apply(gpu, bad)


def radial_case(n=40_000, nx=6, ny=8):
    """Rectilinear axisymmetric box with BOOL_RADIAL_WEIGHTING: CELL_FNUM grows with the radius of the cell row
    (src/initialization.f90:2304-2310)."""
    w = wl.c1_argon_box(n, nx=nx, ny=ny, periodic=False)
    w.cfg.axi = 1
    w.cfg.ymin, w.cfg.ymax = 0.0, 1.0
    w.cfg.zmin, w.cfg.zmax = 0.0, 1.0
    w.cfg.bool_z_periodic = 0
    w.cfg.collision_type = capi.DSMC
    w.cfg.dt = 1.5e-4
    w.cfg.capacity = 4 * n
    w.particles["Z"] = 0.0
    w.particles["DTRIM"] = w.cfg.dt
    j = np.arange(ny)
    ymin, ymax = w.cfg.ymin + j * (w.cfg.ymax - w.cfg.ymin) / ny, w.cfg.ymin + (j + 1) * (w.cfg.ymax - w.cfg.ymin) / ny
    row = w.cfg.fnum * ny * (ymax ** 2 - ymin ** 2) / (w.cfg.ymax ** 2 - w.cfg.ymin ** 2)
    cell_fnum = np.repeat(row, nx)          # index I + NX*(J-1)
    return w, cell_fnum


def test_radial_weighting_clone_and_delete(gpu, oracle_philox):
    """src/timecycle.f90:2110-2128: a particle that moved to a cell row of different weight is deleted with probability
    1 - ratio or cloned floor/rand(ratio - 1) times; one draw per particle (stream kind 8, keyed by the slot).  Same
    survivors and clones on both sides, as a set."""
    w, cell_fnum = radial_case()
    for lib in (oracle_philox, gpu):
        wl.apply(lib, w)
        lib.set_cell_fnum(cell_fnum)
        lib.advect(1)
    a, b = oracle_philox.download_particles(), gpu.download_particles()
    assert len(a) == len(b) and len(a) != len(w.particles)
    key = lambda p: p[np.lexsort((p["VZ"], p["VY"], p["VX"], p["Y"], p["X"], p["IC"]))]
    a, b = key(a), key(b)
    for f in ("IC", "S_ID"):
        assert np.array_equal(a[f], b[f]), f
    for f in ("X", "Y", "VX", "VY", "VZ"):
        assert np.allclose(a[f], b[f], rtol=1e-12, atol=1e-12), f
    # clones exist (several particles with identical state) and so do deletions
    _, counts = np.unique(np.stack([b["X"], b["Y"], b["VX"]]), axis=1, return_counts=True)
    assert (counts > 1).sum() > 50
    # CFNUM = CELL_FNUM(JC) in the collision majorant: same pairs on both sides
    gpu.sort()
    ps = gpu.download_particles()
    oracle_philox.upload_particles(ps)      # the device's sorted order is the order the pairs index into
    co, cg = oracle_philox.collide(2), gpu.collide(2)
    assert co == cg and cg[0] > 0
    util.assert_particles_match(oracle_philox.download_particles(), gpu.download_particles(), rtol=1e-12, atol=1e-9, what="radial DSMC")
    no, mo, ko, io = oracle_philox.checks()
    ng, mg_, kg, ig = gpu.checks()
    assert np.array_equal(no, ng) and np.allclose(ko, kg, rtol=1e-12)
