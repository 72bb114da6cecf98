"""The electrostatic field step on 2-D (plane / axisymmetric triangles) and 3-D (tetrahedra) unstructured meshes
(SURVEY 8(f) row 4): ASSEMBLE_POISSON (src/fields.f90:256-424, 425-637, 639-646, 758-760), SET_WALL_POTENTIAL
(:5609-5711), the right-hand side rows of DEPOSIT_CHARGE (:5538-5547), the linear solve (KSPSolve in the reference,
:5205-5234) and COMPUTE_E_FIELD (:5237-5296).

* `fortran_field_solve` below is an independent plain-numpy reading of the Fortran (per-cell K11..K33 / KIJ, MatSetValue
  ADD_VALUES into a dense matrix, numpy's LAPACK solve) -- not derived from oracle/pantera_oracle.c, which eliminates with
  its own pivoted loop; the two must agree to 1e-10;
* analytic pins: a uniform space charge between two electrodes on a uniform right-triangle mesh is nodally exact (the
  P1 stiffness there is the 5-point stencil); a linear potential is reproduced exactly on ANY mesh (jittered triangles and
  tetrahedra), driven either by two Dirichlet electrodes or by one electrode and a Neumann wall (d phi/dn = WALL_EFIELD);
* the device (Jacobi-preconditioned conjugate gradient, rtol 1e-13) against the oracle (dense elimination): 1e-9 of the
  largest potential, on RF electrodes, Neumann walls, a dielectric volume group (EPS_REL), plane, axisymmetric and 3-D
  meshes, over several timesteps (the previous potential is the starting vector);
* the loop DEPOSIT_CHARGE -> field solve (RHS still on the device) -> ADVECT with the Boris kick on triangles equals the
  oracle's loop.
"""
import numpy as np
import pytest

from pantera_pic_dsmc_b200 import capi, meshgen as mg, workloads as wl

import util

FACE_NODES = {2: [(0, 1), (1, 2), (2, 0)], 3: [(0, 2, 1), (0, 1, 3), (1, 2, 3), (0, 3, 2)]}
DIRICHLET_LIKE = (capi.DIRICHLET_BC, capi.RF_VOLTAGE_BC, capi.DECOUPLED_RF_VOLTAGE_BC)


def fortran_field_solve(mesh, cfg, field_bc, fb, tID, rhs):
    """Plain-numpy evaluation of ASSEMBLE_POISSON + SET_WALL_POTENTIAL + the RHS rows, solved with numpy."""
    d, nn, nc = mesh.dims, mesh.nn, mesh.nc
    xyz = np.asarray(mesh.node_coords)
    cn = np.asarray(mesh.cell_nodes) - 1
    fpg = np.asarray(mesh.cell_faces_pg)
    bcf = np.asarray(mesh.basis_coeffs)
    pot = lambda pg: fb["wall_potential"][pg - 1] + (
        0.0 if field_bc[pg - 1] == capi.DIRICHLET_BC else
        0.5 * fb["wall_rf_potential"][pg - 1] * np.cos(2 * np.pi * fb["rf_frequency"][pg - 1] * tID * cfg.dt))
    is_dir, is_neu = np.zeros(nn, bool), np.zeros(nn, bool)
    dirv, neu = np.zeros(nn), np.zeros(nn)
    used = np.zeros(nn, bool)
    for i in range(nc):
        used[cn[:, i]] = True
        for j in range(d + 1):
            pg = fpg[j, i]
            if pg == -1:
                continue
            for t in FACE_NODES[d][j]:
                v = cn[t, i]
                if field_bc[pg - 1] in DIRICHLET_LIKE:
                    dirv[v], is_dir[v] = pot(pg), True
                if field_bc[pg - 1] == capi.NEUMANN_BC:
                    is_neu[v] = True
    A = np.zeros((nn, nn))
    for i in range(nc):
        eps = 1.0 if mesh.cell_pg[i] == -1 else fb["eps_rel"][mesh.cell_pg[i] - 1]
        v = cn[:, i]
        if d == 2:
            (x1, x2, x3), (y1, y2, y3) = xyz[0, v], xyz[1, v]
            area = 0.5 * abs(x1 * (y2 - y3) + x2 * (y3 - y1) + x3 * (y1 - y2))
            K = np.zeros((3, 3))
            K[0, 0] = 0.25 * ((y2 - y3) ** 2 + (x2 - x3) ** 2) / area * eps
            K[1, 1] = 0.25 * ((y1 - y3) ** 2 + (x1 - x3) ** 2) / area * eps
            K[2, 2] = 0.25 * ((y2 - y1) ** 2 + (x2 - x1) ** 2) / area * eps
            K[0, 1] = K[1, 0] = -0.25 * ((y2 - y3) * (y1 - y3) + (x2 - x3) * (x1 - x3)) / area * eps
            K[1, 2] = K[2, 1] = 0.25 * ((y1 - y3) * (y2 - y1) + (x1 - x3) * (x2 - x1)) / area * eps
            K[0, 2] = K[2, 0] = -0.25 * ((y2 - y3) * (y2 - y1) + (x2 - x3) * (x2 - x1)) / area * eps
            if cfg.axi:
                K = K * (y1 + y2 + y3) / 3.0
        else:
            g = bcf[:3, :, i]                                   # BASIS_COEFFS(1:3, P, I)
            K = mesh.cell_volumes[i] * (g.T @ g) * eps
        for p in range(d + 1):
            if not is_dir[v[p]]:
                A[v[p], v] += K[p]
        for j in range(d + 1):
            pg = fpg[j, i]
            if pg == -1 or field_bc[pg - 1] != capi.NEUMANN_BC:
                continue
            nodes = v[list(FACE_NODES[d][j])]
            P = xyz[:, nodes]
            if d == 2:
                add = fb["wall_efield"][pg - 1] * 0.5 * np.sqrt((P[0, 1] - P[0, 0]) ** 2 + (P[1, 1] - P[1, 0]) ** 2)
            else:
                add = fb["wall_efield"][pg - 1] * 0.5 * np.linalg.norm(np.cross(P[:, 1] - P[:, 0], P[:, 2] - P[:, 0])) / 3.0
            for vv in nodes:
                if not is_dir[vv]:
                    neu[vv] += add
    is_dir |= ~used
    b = np.where(is_dir, dirv, np.where(is_neu, rhs + neu, rhs))
    for v in np.nonzero(is_dir)[0]:
        A[v, :] = 0.0
        A[v, v] = 1.0
    phi = np.linalg.solve(A, b)
    e = np.zeros((3, nc))
    for i in range(nc):
        for k in range(d):
            e[k, i] = -np.dot(phi[cn[:, i]], bcf[k, :, i])
    return phi, e


def full_fb(n, **kw):
    fb = dict(wall_potential=[0.0] * n, wall_rf_potential=[0.0] * n, rf_frequency=[0.0] * n, wall_efield=[0.0] * n, eps_rel=[1.0] * n)
    fb.update(kw)
    return fb


def tri_case(axi=False, jitter=0.3, nx=9, ny=7, dielectric=False):
    """xmin RF electrode, xmax grounded, ymin Neumann wall, ymax nothing (natural); optionally a dielectric block."""
    lo = (0.0, 0.01 if axi else 0.0)
    w = wl.c3_tri_box(nx=nx, ny=ny, n=200, lo=lo, hi=(0.1, 0.06), axi=axi)
    if jitter:
        rng = np.random.default_rng(2)
        X, Y = np.meshgrid(np.linspace(lo[0], 0.1, nx + 1), np.linspace(lo[1], 0.06, ny + 1), indexing="ij")
        X[1:-1, 1:-1] += (rng.random((nx - 1, ny - 1)) - 0.5) * jitter * 0.1 / nx
        Y[1:-1, 1:-1] += (rng.random((nx - 1, ny - 1)) - 0.5) * jitter * (0.06 - lo[1]) / ny
        w.mesh = mg.mapped_quads_to_tris(X, Y, zmin=w.cfg.zmin, zmax=w.cfg.zmax, axi=axi)
    field_bc = [capi.RF_VOLTAGE_BC, capi.DIRICHLET_BC, capi.NEUMANN_BC, capi.NO_BC]
    fb = full_fb(4, wall_potential=[5.0, -3.0, 0.0, 0.0], wall_rf_potential=[120.0, 0.0, 0.0, 0.0],
                 rf_frequency=[13.56e6, 0.0, 0.0, 0.0], wall_efield=[0.0, 0.0, 250.0, 0.0])
    n = 4
    if dielectric:                       # volume group 5 with EPS_REL = 4 over the middle third of the cells
        cpg = np.full(w.mesh.nc, -1, dtype=np.int32)
        cpg[w.mesh.nc // 3: 2 * w.mesh.nc // 3] = 5
        w.mesh.cell_pg = cpg
        field_bc.append(capi.NO_BC)
        fb = {k: v + [4.0 if k == "eps_rel" else 0.0] for k, v in fb.items()}
        n = 5
    w.bcs = dict(particle_bc=[capi.SPECULAR] * n, field_bc=field_bc, wall_temp=[300.0] * n, acc_n=[1.0] * n, acc_t=[1.0] * n)
    w.particles = w.particles[:0]
    w.cfg.dt = 1e-9
    return w, field_bc, fb


def tet_case(jitter=0.15):
    """xmin Dirichlet, xmax RF electrode, ymin Neumann, the rest natural."""
    w = wl.c5_channel(4, 3, 3, ppc=1, jitter=jitter)
    field_bc = [capi.DIRICHLET_BC, capi.RF_VOLTAGE_BC, capi.NEUMANN_BC, capi.NO_BC, capi.NO_BC, capi.NO_BC]
    fb = full_fb(6, wall_potential=[2.0, 0.0, 0.0, 0.0, 0.0, 0.0], wall_rf_potential=[0.0, 60.0, 0.0, 0.0, 0.0, 0.0],
                 rf_frequency=[0.0, 2.0e6, 0.0, 0.0, 0.0, 0.0], wall_efield=[0.0, 0.0, -4.0e3, 0.0, 0.0, 0.0])
    w.bcs = dict(particle_bc=[capi.SPECULAR] * 6, field_bc=field_bc, wall_temp=[300.0] * 6, acc_n=[1.0] * 6, acc_t=[1.0] * 6)
    w.particles = w.particles[:0]
    w.cfg.dt = 1e-8
    return w, field_bc, fb


CASES = {
    "tris": lambda: tri_case(),
    "tris-dielectric": lambda: tri_case(dielectric=True),
    "tris-axi": lambda: tri_case(axi=True),
    "tets": lambda: tet_case(),
}


def rhs_for(mesh, seed=3):
    return np.random.default_rng(seed).standard_normal(mesh.nn) * 5.0


def close(a, b, tol):
    return np.allclose(a, b, rtol=0.0, atol=tol * max(np.abs(a).max(), 1e-300))


@pytest.mark.parametrize("case", list(CASES))
def test_oracle_field_solve_equals_the_plain_numpy_reading_of_the_fortran(oracle, case):
    w, field_bc, fb = CASES[case]()
    wl.apply(oracle, w)
    oracle.set_field_bcs(**fb)
    rhs = rhs_for(w.mesh)
    for t in (0, 17, 350):
        phi, e, _, _ = oracle.field_solve(t, w.mesh.nn, w.mesh.nc, rhs=rhs)
        pr, er = fortran_field_solve(w.mesh, w.cfg, field_bc, fb, t, rhs)
        assert close(pr, phi, 1e-10) and close(er, e, 1e-9), (case, t)
    # the RF electrode follows WALL_POTENTIAL + 0.5*WALL_RF_POTENTIAL*cos(2 pi f tID DT) and the Dirichlet one its value
    k = field_bc.index(capi.RF_VOLTAGE_BC)
    on_rf = np.unique((np.asarray(w.mesh.cell_nodes) - 1)[list(FACE_NODES[w.mesh.dims][0])][:, np.asarray(w.mesh.cell_faces_pg)[0] == k + 1])
    if on_rf.size:
        expect = fb["wall_potential"][k] + 0.5 * fb["wall_rf_potential"][k] * np.cos(2 * np.pi * fb["rf_frequency"][k] * 350 * w.cfg.dt)
        assert np.allclose(phi[on_rf], expect, rtol=1e-14)


def uniform_load(mesh, rho):
    """RHS_i = integral of rho psi_i over the mesh (P1: a third / a quarter of every adjacent cell)."""
    cn = np.asarray(mesh.cell_nodes) - 1
    size = np.asarray(mesh.cell_volumes)
    if mesh.dims == 2:
        xyz = np.asarray(mesh.node_coords)
        P = xyz[:, cn]                                   # (3, 3 nodes, nc)
        size = 0.5 * np.abs((P[0, 1] - P[0, 0]) * (P[1, 2] - P[1, 0]) - (P[0, 2] - P[0, 0]) * (P[1, 1] - P[1, 0]))
    rhs = np.zeros(mesh.nn)
    for p in range(mesh.dims + 1):
        np.add.at(rhs, cn[p], rho * size / (mesh.dims + 1))
    return rhs


def check_parabola(lib, rtol=0.0):
    """-phi'' = rho between phi(0) = 0 and phi(L) = V on a uniform right-triangle mesh: nodally exact."""
    w = wl.c3_tri_box(nx=12, ny=5, n=10, lo=(0.0, 0.0), hi=(0.1, 0.05))
    w.bcs = dict(particle_bc=[capi.SPECULAR] * 4, field_bc=[capi.DIRICHLET_BC, capi.DIRICHLET_BC, capi.NO_BC, capi.NO_BC],
                 wall_temp=[300.0] * 4, acc_n=[1.0] * 4, acc_t=[1.0] * 4)
    w.particles = w.particles[:0]
    wl.apply(lib, w)
    V, rho, L = 37.0, 5.0e4, 0.1
    lib.set_field_bcs(wall_potential=[0.0, V, 0.0, 0.0])
    phi, e, it, rel = lib.field_solve(0, w.mesh.nn, w.mesh.nc, rhs=uniform_load(w.mesh, rho), rtol=rtol, max_iter=5000)
    x = np.asarray(w.mesh.node_coords)[0]
    exact = V * x / L + 0.5 * rho * (x * L - x * x)
    assert close(exact, phi, 1e-10), np.abs(phi - exact).max()
    return it, rel


def check_linear(lib, kind, neumann, rtol=0.0):
    """A linear potential is in the P1 space of any mesh: two electrodes, or one electrode and d phi/dn = g on the other
    side (NEUMANN(:) = g * edge/2 or g * face/3 per node is the exact boundary integral of g psi_i)."""
    if kind == "tris":
        w, _, _ = tri_case(jitter=0.35)
        L, nb = 0.1, 4
    else:
        w, _, _ = tet_case(jitter=0.2)
        L, nb = w.cfg.xmax, 6
    g = 800.0
    field_bc = [capi.DIRICHLET_BC, capi.NEUMANN_BC if neumann else capi.DIRICHLET_BC] + [capi.NO_BC] * (nb - 2)
    w.bcs["field_bc"] = field_bc
    wl.apply(lib, w)
    lib.set_field_bcs(wall_potential=[0.0, 0.0 if neumann else g * L] + [0.0] * (nb - 2), wall_efield=[0.0, g if neumann else 0.0] + [0.0] * (nb - 2))
    phi, e, _, _ = lib.field_solve(0, w.mesh.nn, w.mesh.nc, rhs=np.zeros(w.mesh.nn), rtol=rtol, max_iter=5000)
    x = np.asarray(w.mesh.node_coords)[0]
    assert close(g * x, phi, 1e-10), np.abs(phi - g * x).max()
    assert close(np.full(w.mesh.nc, -g), e[0], 1e-9) and np.abs(e[1:]).max() < 1e-8 * g


def test_oracle_parabola_on_uniform_triangles_is_nodally_exact(oracle):
    check_parabola(oracle)


@pytest.mark.parametrize("kind", ["tris", "tets"])
@pytest.mark.parametrize("neumann", [False, True])
def test_oracle_reproduces_a_linear_potential_on_jittered_meshes(oracle, kind, neumann):
    check_linear(oracle, kind, neumann)


# ---- device ------------------------------------------------------------------------------------------------------
@pytest.mark.gpu
def test_device_parabola_on_uniform_triangles_is_nodally_exact(gpu):
    it, rel = check_parabola(gpu, rtol=1e-13)
    assert 0 < it < 1000 and rel <= 1e-13


@pytest.mark.gpu
@pytest.mark.parametrize("kind", ["tris", "tets"])
@pytest.mark.parametrize("neumann", [False, True])
def test_device_reproduces_a_linear_potential_on_jittered_meshes(gpu, kind, neumann):
    check_linear(gpu, kind, neumann, rtol=1e-13)


@pytest.mark.gpu
@pytest.mark.parametrize("coop", ["1", "0"])          # one cooperative launch / one kernel per half iteration
@pytest.mark.parametrize("case", list(CASES))
def test_device_field_solve_matches_the_oracle(gpu, oracle, case, coop, monkeypatch):
    monkeypatch.setenv("PANTERA_GPU_CG_COOP", coop)
    w, field_bc, fb = CASES[case]()
    rhs = rhs_for(w.mesh)
    for lib in (oracle, gpu):
        wl.apply(lib, w)
        lib.set_field_bcs(**fb)
    iters = []
    for t in (0, 17, 18, 350):            # (17 -> 18: the warm start is nearly the answer)
        po, eo, _, _ = oracle.field_solve(t, w.mesh.nn, w.mesh.nc, rhs=rhs)
        pg, eg, it, rel = gpu.field_solve(t, w.mesh.nn, w.mesh.nc, rhs=rhs, rtol=1e-13, max_iter=5000)
        assert rel <= 1e-13 and it < 5000, (case, t, it, rel)
        assert close(po, pg, 1e-9) and close(eo, eg, 1e-8), (case, t, np.abs(po - pg).max())
        iters.append(it)
    assert iters[2] <= iters[1]
    # PETSc's default tolerance: fewer iterations, residual as asked
    gpu.set_field_bcs(**fb)               # (drops the warm start)
    _, _, it5, rel5 = gpu.field_solve(350, w.mesh.nn, w.mesh.nc, rhs=rhs)
    assert rel5 <= 1e-5 and 0 < it5 < max(iters[0], 2)
    # a second solve of the same system starts from the answer
    _, _, it0, _ = gpu.field_solve(350, w.mesh.nn, w.mesh.nc, rhs=rhs)
    assert it0 == 0


@pytest.mark.gpu
def test_device_field_solve_is_reproducible_bit_for_bit(gpu):
    """Every reduction of the solver runs in a fixed order: the same system gives the same bits."""
    w, field_bc, fb = tri_case(nx=40, ny=30)
    rhs = rhs_for(w.mesh)
    outs = []
    for _ in range(2):
        wl.apply(gpu, w)
        gpu.set_field_bcs(**fb)
        outs.append(gpu.field_solve(5, w.mesh.nn, w.mesh.nc, rhs=rhs, rtol=1e-12, max_iter=5000))
    assert outs[0][2] == outs[1][2] and np.array_equal(outs[0][0], outs[1][0]) and np.array_equal(outs[0][1], outs[1][1])


@pytest.mark.gpu
def test_device_field_solve_errors(gpu):
    w, field_bc, fb = tri_case()
    w.bcs["field_bc"] = [capi.NEUMANN_BC, capi.NO_BC, capi.NO_BC, capi.NO_BC]
    wl.apply(gpu, w)
    gpu.set_field_bcs(**fb)
    with pytest.raises(capi.PanteraError, match="singular"):
        gpu.field_solve(0, w.mesh.nn, w.mesh.nc, rhs=rhs_for(w.mesh))
    w, field_bc, fb = tri_case()
    wl.apply(gpu, w)
    with pytest.raises(capi.PanteraError, match="RHS"):
        gpu.field_solve(0, w.mesh.nn, w.mesh.nc)          # nothing deposited yet


@pytest.mark.gpu
def test_pic_loop_on_triangles_matches_oracle(gpu, oracle_philox):
    """DEPOSIT_CHARGE -> Poisson (CG on the device, RHS never leaves it) -> E -> ADVECT with the Boris kick, 4 steps."""
    w, field_bc, fb = tri_case(jitter=0.25, nx=10, ny=8)
    src = wl.c3_tri_box(nx=10, ny=8, n=20_000, lo=(0.0, 0.0), hi=(0.1, 0.06), T0=20000.0, u_drift=(0.0, 0.0, 0.0))
    sp = wl.load_species("ar", ["Ar+", "e"])
    rng = np.random.default_rng(12)
    p = src.particles
    n = len(p)
    p["S_ID"] = rng.integers(1, 3, n)
    v = wl.maxwellian(rng, n, sp.mass[p["S_ID"] - 1], 20000.0)
    p["VX"], p["VY"], p["VZ"] = v[:, 0], v[:, 1], v[:, 2]
    p["EROT"] = 0.0
    p["EVIB"] = 0.0
    w.species = sp
    w.cfg.pic_type = capi.PIC_EXPLICIT
    w.cfg.dt = 2e-10
    w.cfg.fnum = 2.0e4
    p["DTRIM"] = w.cfg.dt
    w.bcs["particle_bc"] = [capi.VACUUM, capi.VACUUM, capi.SPECULAR, capi.SPECULAR]
    w.particles = p
    w.cfg.capacity = 2 * n
    # the particles were placed on the unjittered mesh: locate them on this one (the re-binning is tested on its own)
    util.setup_pair(gpu, oracle_philox, w)
    assert oracle_philox.reassign_particles_to_cells() == gpu.reassign_particles_to_cells()
    for lib in (gpu, oracle_philox):
        lib.set_field_bcs(**fb)
    for t in range(1, 5):
        rg = gpu.deposit_charge(w.mesh.nn)
        ro = oracle_philox.deposit_charge(w.mesh.nn)
        assert np.allclose(rg, ro, rtol=1e-10, atol=1e-12 * np.abs(ro).max())
        pg, eg, it, rel = gpu.field_solve(t, w.mesh.nn, w.mesh.nc, rtol=1e-13, max_iter=5000)
        po, eo, _, _ = oracle_philox.field_solve(t, w.mesh.nn, w.mesh.nc, rhs=ro)
        assert rel <= 1e-13 and close(po, pg, 1e-9) and close(eo, eg, 1e-8)
        gpu.advect(t)
        oracle_philox.advect(t)
        a, b = util.by_id(oracle_philox.download_particles()), util.by_id(gpu.download_particles())
        assert np.array_equal(a["ID"], b["ID"])
        util.assert_particles_match(a, b, rtol=1e-7, atol=1e-12, what=f"2-D PIC step {t}")
        oracle_philox.upload_particles(b)      # resync (deposit order and CG tolerance differ at the 1e-9 level)
        gpu.upload_particles(b)
    assert np.abs(eg).max() > 0
