"""The electrostatic field step on the device: SET_WALL_POTENTIAL + the 1-D Poisson system of
ASSEMBLE_POISSON + COMPUTE_E_FIELD (src/fields.f90:137-253, 758-760, 5237-5296, 5538-5607).

* analytic pin: uniform space charge between a grounded and a biased electrode -- the P1 finite-element solution is
  nodally exact, phi(x) = V x/L + (rho/2)(x L - x^2) for -phi'' = rho;
* the oracle assembles the matrix with the reference's MatSetValue rules and solves it densely (partial pivoting);
  the device eliminates along the chain of nodes; both must agree to 1e-10;
* an RF electrode follows WALL_POTENTIAL + 0.5*WALL_RF_POTENTIAL*cos(2 pi f tID DT);
* end to end: the C4 loop DEPOSIT_CHARGE -> field solve -> ADVECT (Boris kick) -> MCC on the device equals the oracle's
  loop value for value.
"""
import numpy as np
import pytest

from pantera_pic_dsmc_b200 import capi, workloads as wl

import util


def ccp(ncells=48, **kw):
    w = wl.c4_ccp_1d(ncells, 8_000, **kw)
    return w


def uniform_rhs(mesh, rho):
    """RHS_i = integral of rho * psi_i (P1 hats on the line)"""
    x = np.asarray(mesh.node_coords)[0]
    cn = np.asarray(mesh.cell_nodes) - 1
    rhs = np.zeros(mesh.nn)
    for c in range(mesh.nc):
        L = abs(x[cn[1, c]] - x[cn[0, c]])
        rhs[cn[0, c]] += 0.5 * rho * L
        rhs[cn[1, c]] += 0.5 * rho * L
    return rhs


def check_analytic(lib, nonuniform):
    w = ccp(40)
    m = w.mesh
    if nonuniform:  # stretch the mesh: nodes at L*s^1.7 (and rebuild the P1 basis the same way meshgen does)
        from pantera_pic_dsmc_b200 import meshgen as mg
        s = np.linspace(0, 1, m.nn) ** 1.7
        m2 = mg.line_1d(m.nc, 0.0, 0.0254)
        xs = 0.0254 * s
        coords = np.asarray(m2.node_coords).copy()
        coords[0, :] = xs
        m2.node_coords = np.asfortranarray(coords)
        cn = np.asarray(m2.cell_nodes) - 1
        bc = np.zeros((2, 2, m2.nc))
        for c in range(m2.nc):
            x1, x2 = xs[cn[0, c]], xs[cn[1, c]]
            bc[0, 0, c], bc[1, 0, c] = 1.0 / (x1 - x2), -x2 / (x1 - x2)
            bc[0, 1, c], bc[1, 1, c] = 1.0 / (x2 - x1), -x1 / (x2 - x1)
        m2.basis_coeffs = np.asfortranarray(bc)
        m2.cell_volumes = np.abs(xs[cn[1]] - xs[cn[0]])
        w.mesh = m = m2
    w.particles = w.particles[:0]
    wl.apply(lib, w)
    V, rho, L = 37.0, 5.0e4, 0.0254
    lib.set_field_bcs(wall_potential=[0.0, V])
    rhs = uniform_rhs(m, rho)
    phi, e = lib.field_solve_1d(3, m.nn, m.nc, rhs=rhs)
    x = np.asarray(m.node_coords)[0]
    exact = V * x / L + 0.5 * rho * (x * L - x * x)
    assert np.allclose(phi, exact, rtol=1e-10, atol=1e-10 * np.abs(exact).max())
    cn = np.asarray(m.cell_nodes) - 1
    e_exact = -(exact[cn[1]] - exact[cn[0]]) / (x[cn[1]] - x[cn[0]])
    assert np.allclose(e[0], e_exact, rtol=1e-9, atol=1e-9 * np.abs(e_exact).max()) and not e[1:].any()


@pytest.mark.parametrize("nonuniform", [False, True])
def test_oracle_poisson_1d_is_nodally_exact(oracle, nonuniform):
    check_analytic(oracle, nonuniform)


@pytest.mark.gpu
@pytest.mark.parametrize("nonuniform", [False, True])
def test_device_poisson_1d_is_nodally_exact(gpu, nonuniform):
    check_analytic(gpu, nonuniform)


@pytest.mark.gpu
def test_device_field_step_matches_oracle_with_rf_electrode_and_neumann_wall(gpu, oracle):
    w = ccp(64, field_bc=(capi.RF_VOLTAGE_BC, capi.NEUMANN_BC))
    rng = np.random.default_rng(3)
    rhs = rng.standard_normal(w.mesh.nn) * 10.0
    outs = []
    for lib in (oracle, gpu):
        wl.apply(lib, w)
        lib.set_field_bcs(wall_potential=[5.0, 0.0], wall_rf_potential=[200.0, 0.0], rf_frequency=[13.56e6, 0.0], wall_efield=[0.0, 30.0])
        outs.append([lib.field_solve_1d(t, w.mesh.nn, w.mesh.nc, rhs=rhs) for t in (0, 1234, 5000)])
    for (po, eo), (pg, eg) in zip(*outs):
        assert np.allclose(po, pg, rtol=1e-10, atol=1e-10 * np.abs(po).max())
        assert np.allclose(eo, eg, rtol=1e-9, atol=1e-9 * np.abs(eo).max())
    # the electrode follows the RF drive
    v0 = outs[1][0][0][0]
    assert v0 == pytest.approx(5.0 + 100.0)
    t = 1234
    assert outs[1][1][0][0] == pytest.approx(5.0 + 100.0 * np.cos(2 * np.pi * 13.56e6 * t * w.cfg.dt))


@pytest.mark.gpu
def test_compute_e_field_from_host_potential_on_triangles(gpu, oracle):
    w = wl.c3_tri_box(nx=7, ny=5, n=100)
    rng = np.random.default_rng(5)
    phi = rng.standard_normal(w.mesh.nn)
    for lib in (oracle, gpu):
        wl.apply(lib, w)
    eo, eg = oracle.compute_e_field(phi, w.mesh.nc), gpu.compute_e_field(phi, w.mesh.nc)
    assert np.allclose(eo, eg, rtol=1e-13, atol=1e-13 * np.abs(eo).max()) and np.abs(eo[:2]).max() > 0 and not eo[2].any()


@pytest.mark.gpu
def test_ccp_loop_on_device_matches_oracle(gpu, oracle_philox):
    """DEPOSIT_CHARGE -> Poisson -> E -> ADVECT -> MCC, five steps, particles matched by ID."""
    w = ccp(48, wall_bc=("vacuum", "vacuum"), field_bc=(capi.RF_VOLTAGE_BC, capi.DIRICHLET_BC))
    w.cfg.dt = 2e-10
    w.particles["DTRIM"] = w.cfg.dt
    w.cfg.capacity = 4 * len(w.particles)
    w.efield = None
    util.setup_pair(gpu, oracle_philox, w)
    for lib in (gpu, oracle_philox):
        lib.set_field_bcs(wall_potential=[0.0, 0.0], wall_rf_potential=[80.0, 0.0], rf_frequency=[13.56e6, 0.0])
    for t in range(1, 6):
        rg = gpu.deposit_charge(w.mesh.nn)
        ro = oracle_philox.deposit_charge(w.mesh.nn)
        assert np.allclose(rg, ro, rtol=1e-10, atol=1e-12 * np.abs(ro).max())
        pg, eg = gpu.field_solve_1d(t, w.mesh.nn, w.mesh.nc)            # the RHS still on the device
        po, eo = oracle_philox.field_solve_1d(t, w.mesh.nn, w.mesh.nc, rhs=ro)
        assert np.allclose(pg, po, rtol=1e-9, atol=1e-9 * np.abs(po).max())
        gpu.advect(t)
        oracle_philox.advect(t)
        a, b = util.by_id(oracle_philox.download_particles()), util.by_id(gpu.download_particles())
        assert np.array_equal(a["ID"], b["ID"])
        util.assert_particles_match(a, b, rtol=1e-8, atol=1e-12, what=f"CCP step {t}")
        oracle_philox.upload_particles(b)      # resync (the two deposits differ in summation order by 1e-12)
        gpu.upload_particles(b)
    assert len(b) < len(w.particles)          # some electrons reached the electrodes
