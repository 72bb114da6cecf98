"""Independent pins: vectors evaluated in plain Python from the FORTRAN source (tests/golden/fortran_pins.py ->
tests/golden/fortran_pins.json), replayed through the oracle (scripted draws, rng_mode 2) and, for the deterministic
events, through the CUDA library.  Reference lines covered: see the header of tests/golden/fortran_pins.py and
DESIGN.md section 5.
"""
import ctypes as C
import json
import os
import types

import numpy as np
import pytest

from pantera_pic_dsmc_b200 import capi, datafiles as df, workloads as wl

import util

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
PINS = json.load(open(os.path.join(GOLD, "fortran_pins.json")))
fh = float.fromhex


def script(orc, draws):
    a = np.ascontiguousarray(np.asarray(draws, dtype=np.float64))
    orc._call("set_rng_mode", C.c_int(2))
    orc._call("set_rng_script", a.ctypes.data_as(C.c_void_p), C.c_int64(len(a)))


def script_used(orc):
    f = orc.lib.po_rng_script_used
    f.restype = C.c_int64
    return int(f(orc.ctx))


def pin_tet_mesh():
    m = PINS["tet_mesh"]
    nn, nc = len(m["nodes"]), len(m["cells"])
    o = types.SimpleNamespace(dims=3, nn=nn, nc=nc)
    o.node_coords = np.asfortranarray(np.array(m["nodes"], dtype=np.float64).T)                  # (3, nn)
    o.cell_nodes = np.asfortranarray(np.array(m["cells"], dtype=np.int32).T)                     # (4, nc)
    o.cell_neighbors = np.asfortranarray(np.array(m["neighbors"], dtype=np.int32).T)
    o.cell_faces_pg = np.asfortranarray(np.array(m["faces_pg"], dtype=np.int32).T)
    for k in ("normal", "tang1", "tang2"):                                                       # (3, 4, nc)
        setattr(o, "face_" + k, np.asfortranarray(np.array(m[k], dtype=np.float64).transpose(2, 1, 0)))
    o.cell_volumes = np.array(m["vol"], dtype=np.float64)
    o.cell_pg = None
    o.basis_coeffs = None
    o.cell_edges_len = None
    o.cell_procs = None
    o.face_boundary_index = None
    o.nboundcells = 0
    o.boundary_area = None
    return o


def tet_workload(cases, dt, bc_kind=capi.SPECULAR, wall=None):
    sp = wl.load_species("air", ["N2", "O2"])
    cfg = capi.make_config(dims=3, grid_type=capi.UNSTRUCTURED, nx=1, ny=1, xmin=0.0, xmax=2.0, ymin=0.0, ymax=2.0,
                           zmin=0.0, zmax=2.0, dt=dt, fnum=1.0, collision_type=capi.NO_COLL, dump_bound_avg_every=0)
    p = wl.new_particles(len(cases))
    for k, c in enumerate(cases):
        q = c["inp"]
        for f in ("x", "y", "z", "vx", "vy", "vz", "erot", "evib"):
            p[f.upper()][k] = fh(q[f])
        p["IC"][k] = q["ic"]
        p["S_ID"][k] = sp.id(q["sid"])
    p["DTRIM"] = dt
    kinds = [capi.SPECULAR] * 4
    temp, an, at = [300.0] * 4, [1.0] * 4, [1.0] * 4
    if wall is not None:
        kinds[0] = bc_kind
        temp[0], an[0], at[0] = wall["temp"], wall["acc_n"], wall["acc_t"]
    bcs = dict(particle_bc=kinds, wall_temp=temp, acc_n=an, acc_t=at)
    return wl.Workload("pin tets", cfg, sp, p, mesh=pin_tet_mesh(), bcs=bcs)


def check_walk(lib, exact):
    t = PINS["tet_walk"]
    w = tet_workload(t["cases"], fh(t["dt"]))
    w.cfg.track_id = 1
    wl.apply(lib, w)
    lib.advect(1)
    out = util.by_id(lib.download_particles())
    for k, c in enumerate(t["cases"]):
        e = c["out"]
        assert out["IC"][k] == e["ic"], (k, out["IC"][k], e["ic"], c["trace"])
        for f in ("x", "y", "z", "vx", "vy", "vz"):
            got, want = float(out[f.upper()][k]), fh(e[f])
            if exact:
                assert got == want, (k, f, got.hex(), e[f])
            else:
                assert got == pytest.approx(want, rel=1e-12, abs=1e-300), (k, f)


def test_tet_walk_oracle_matches_fortran_reading(oracle):
    """Face loop with the '>=' tie (last face wins), hops, specular reflection: bit-identical."""
    assert PINS["tet_walk"]["cases"][0]["trace"][0] == [1, 4]
    check_walk(oracle, exact=True)


@pytest.mark.gpu
def test_tet_walk_device_matches_fortran_reading(gpu):
    check_walk(gpu, exact=True)


def axi_workload():
    w = wl.c3_tri_box(nx=1, ny=1, n=4, axi=True, lo=(0.0, 0.01), hi=(0.1, 0.06), dt=1e-6, wall_bc=("specular",) * 4)
    m = w.mesh
    xy = np.asarray(m.node_coords)[:2]
    tri = np.asarray(m.cell_nodes) - 1

    def cell_of(x, y):
        for c in range(m.nc):
            a, b, cc = (xy[:, tri[j, c]] for j in range(3))
            d = (b[0] - a[0]) * (cc[1] - a[1]) - (cc[0] - a[0]) * (b[1] - a[1])
            l1 = ((b[0] - x) * (cc[1] - y) - (cc[0] - x) * (b[1] - y)) / d
            l2 = ((cc[0] - x) * (a[1] - y) - (a[0] - x) * (cc[1] - y)) / d
            if l1 >= 0 and l2 >= 0 and l1 + l2 <= 1:
                return c + 1
        raise AssertionError("point outside the mesh")
    return w, cell_of


def check_axi(lib):
    for k, c in enumerate(PINS["axi_move"]):
        w, cell_of = axi_workload()
        q = c["inp"]
        p = wl.new_particles(1)
        for f in ("x", "y", "z", "vx", "vy", "vz"):
            p[f.upper()][0] = fh(q[f])
        p["S_ID"] = 1
        p["IC"] = cell_of(fh(q["x"]), fh(q["y"]))
        w.cfg.dt = fh(c["time"])
        p["DTRIM"] = w.cfg.dt
        w.cfg.collision_type = capi.NO_COLL
        w.particles = p
        wl.apply(lib, w)
        lib.advect(1)
        out = lib.download_particles()
        assert len(out) == 1 and out["IC"][0] == p["IC"][0]
        for f in ("x", "y", "z", "vx", "vy", "vz"):
            assert float(out[f.upper()][0]) == fh(c["out"][f]), (k, f, float(out[f.upper()][0]).hex(), c["out"][f])


def test_axisymmetric_move_oracle_matches_fortran_reading(oracle):
    check_axi(oracle)


@pytest.mark.gpu
def test_axisymmetric_move_device_matches_fortran_reading(gpu):
    check_axi(gpu)


def test_cll_reflection_oracle_matches_fortran_reading(oracle):
    t = PINS["cll_reflect"]
    w = tet_workload([t], fh(t["dt"]), bc_kind=capi.CLL, wall=t["wall"])
    wl.apply(oracle, w)
    script(oracle, t["draws"])
    oracle.advect(1)
    assert script_used(oracle) == len(t["draws"])
    out = oracle.download_particles()
    e = t["out"]
    assert out["IC"][0] == e["ic"]
    for f in ("x", "y", "z", "vx", "vy", "vz", "erot", "evib"):
        # NORM2 of the tangent is evaluated by its definition in the pin and by gfortran's scaled loop in the oracle
        assert float(out[f.upper()][0]) == pytest.approx(fh(e[f]), rel=1e-13, abs=1e-300), f


def check_wall_react(lib, scripted):
    """WALL_REACT (src/fields.f90:4302-4341) on a reacting specular wall: conversion with the mass-ratio velocity
    scaling, removal with the renormalised probability, and the missing EXIT (a converted particle can still be removed
    by a later reaction of its ORIGINAL species)."""
    t = PINS["wall_react"]
    for c in t["cases"]:
        w = tet_workload([c], fh(t["dt"]))
        w.bcs["react"] = [1, 0, 0, 0]
        sp = w.species
        w.wall_reactions = dict(r_sp=[sp.id(r["r"]) for r in t["reactions"]], p1_sp=[sp.id(r["p1"]) if r["p1"] else -1 for r in t["reactions"]],
                                nprod=[r["nprod"] for r in t["reactions"]], prob=[r["prob"] for r in t["reactions"]])
        wl.apply(lib, w)
        script(lib, c["draws"])
        lib.advect(1)
        assert script_used(lib) == len(c["draws"])
        out = lib.download_particles()
        if c["removed"]:
            assert len(out) == 0
            continue
        e = c["out"]
        assert len(out) == 1 and out["S_ID"][0] == sp.id(e["sid"]) and out["IC"][0] == e["ic"]
        for f in ("x", "y", "z", "vx", "vy", "vz", "erot", "evib"):
            assert float(out[f.upper()][0]) == fh(e[f]), (f, float(out[f.upper()][0]).hex(), e[f])


def test_wall_react_oracle_matches_fortran_reading(oracle):
    check_wall_react(oracle, True)


def test_diffuse_reflection_oracle_matches_fortran_reading(oracle):
    """Maxwell wall (src/timecycle.f90:1445-1476) on scripted draws: MAXWELL with a redrawn R1 (src/tools.f90:130-187),
    INTERNAL_ENERGY twice, FLX with one rejected pair (src/tools.f90:1150-1186); FACE_TANG1/2 of the pin mesh."""
    t = PINS["diffuse_reflect"]
    w = tet_workload([t], fh(t["dt"]), bc_kind=capi.DIFFUSE, wall=t["wall"])
    wl.apply(oracle, w)
    script(oracle, t["draws"])
    oracle.advect(1)
    assert script_used(oracle) == len(t["draws"])
    out = oracle.download_particles()
    e = t["out"]
    assert out["IC"][0] == e["ic"]
    for f in ("x", "y", "z", "vx", "vy", "vz", "erot", "evib"):
        assert float(out[f.upper()][0]) == pytest.approx(fh(e[f]), rel=1e-14, abs=1e-300), f


def collision_setup(orc, t, with_reactions):
    names = t["species"]
    sp = wl.load_species("air", names)
    assert sp.sigmamax == pytest.approx(fh(t["sigmamax"]), rel=1e-14)
    sp.sigmamax = fh(t["sigmamax"])
    cfg = capi.make_config(dims=2, grid_type=capi.RECTILINEAR_UNIFORM, nx=1, ny=1, xmin=0.0, xmax=1.0, ymin=0.0, ymax=1.0,
                           zmin=-0.5, zmax=0.5, dt=fh(t["dt"]), fnum=fh(t["fnum"]), collision_type=capi.DSMC,
                           bool_periodic=[1] * 4, bool_z_periodic=1, capacity=64)
    p = wl.new_particles(len(t["inp"]))
    for k, q in enumerate(t["inp"]):
        p["X"][k], p["Y"][k], p["Z"][k] = q["pos"]
        p["VX"][k], p["VY"][k], p["VZ"][k] = (fh(c) for c in q["v"])
        p["EROT"][k], p["EVIB"][k] = fh(q["erot"]), fh(q["evib"])
        p["S_ID"][k] = sp.id(q["sid"])
    p["IC"] = 1
    p["DTRIM"] = cfg.dt
    reactions = df.read_reactions(os.path.join(wl.DATA_DIR, "air.tce"), sp) if with_reactions else []
    w = wl.Workload("pin pair", cfg, sp, p, reactions=reactions)
    wl.apply(orc, w)
    return sp, p


def check_pair(out, t, sp, rel):
    for k, e in enumerate(t["out"]):
        assert out["S_ID"][k] == sp.id(e["sid"]), (k, out["S_ID"][k], e["sid"])
        got = [out["VX"][k], out["VY"][k], out["VZ"][k], out["EROT"][k], out["EVIB"][k]]
        want = [fh(c) for c in e["v"]] + [fh(e["erot"]), fh(e["evib"])]
        assert got == pytest.approx(want, rel=rel, abs=1e-300), (k, got, want)


def test_vss_pair_with_borgnakke_larsen_oracle_matches_fortran_reading(oracle):
    """Selection with a repeated index, acceptance, rotational exchange on particle 1 with one rejection in
    COLL_INTERNAL_ENERGY, VSS deflection, isotropic redirection."""
    t = PINS["vss_pair_bl"]
    sp, p = collision_setup(oracle, t, with_reactions=False)
    script(oracle, t["draws"])
    nc, nr = oracle.collide(1)
    assert (nc, nr) == (1, 0) and script_used(oracle) == t["used"]
    check_pair(oracle.download_particles(), t, sp, rel=1e-14)


def test_tce_three_product_dissociation_oracle_matches_fortran_reading(oracle):
    """N2 + O2 --> N2 + O + O: TCE probability, energy redistribution over the products' modes, two HS_SCATTERs, the
    third product created at the position of the second reactant."""
    t = PINS["tce_3prod"]
    sp, p = collision_setup(oracle, t, with_reactions=True)
    script(oracle, t["draws"])
    nc, nr = oracle.collide(1)
    assert (nc, nr) == (1, 1) and script_used(oracle) == t["used"]
    out = oracle.download_particles()
    assert len(out) == 3
    check_pair(out, t, sp, rel=1e-13)
    c = t["created"][0]
    assert out["S_ID"][2] == sp.id(c["sid"]) and out["IC"][2] == 1
    assert [out["X"][2], out["Y"][2], out["Z"][2]] == c["pos"]
    got = [out["VX"][2], out["VY"][2], out["VZ"][2], out["EROT"][2], out["EVIB"][2]]
    assert got == pytest.approx([fh(x) for x in c["v"]] + [fh(c["erot"]), fh(c["evib"])], rel=1e-13, abs=1e-300)


def test_tce_coefficients_match_fortran_reading():
    """INITREACTIONS (src/initialization.f90:2681-2724): C1, C2, C3 of every reaction of data/air.tce."""
    sp = wl.load_species("air", None)
    reac = df.read_reactions(os.path.join(wl.DATA_DIR, "air.tce"), sp)
    seen = 0
    for r in reac:
        names = [sp.names[r["r1"] - 1], sp.names[r["r2"] - 1]]
        prods = [sp.names[r[k] - 1] for k in ("p1", "p2", "p3") if r.get(k, -1) > 0]
        key = " + ".join(names) + " --> " + " + ".join(prods)
        assert key in PINS["tce_coeffs"], key
        c1, c2, c3, ea = (fh(x) for x in PINS["tce_coeffs"][key])
        assert [r["c1"], r["c2"], r["c3"], r["ea"]] == pytest.approx([c1, c2, c3, ea], rel=1e-13)
        seen += 1
    assert seen == len(PINS["tce_coeffs"]) and seen >= 10


# ------------------------------------------------------------------------------------------ radial weighting
def radial_reading(p0, cell_fnum, cfg, draws):
    """src/timecycle.f90:2100-2131 read in plain Python: the reverse sweep IP = NP..1 over the moved particles,
    CELL_FROM_POSITION (src/grid_and_partition.f90:34-60, uniform grid), one draw per particle that changed to a cell
    of different weight, REMOVE_PARTICLE_ARRAY = overwrite with the last element, clones appended at the end.
    Returns the IDs in final array order and the number of draws consumed."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("fortran_pins_gen", os.path.join(GOLD, "fortran_pins.py"))
    gen = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(gen)
    dx = (cfg.xmax - cfg.xmin) / cfg.nx
    dy = (cfg.ymax - cfg.ymin) / cfg.ny
    arr = []  # the particle array after the move: (id, old_ic, x, y)
    for q in p0:
        m = gen.move_particle_axi(dict(x=float(q["X"]), y=float(q["Y"]), z=float(q["Z"]), vx=float(q["VX"]), vy=float(q["VY"]),
                                       vz=float(q["VZ"])), cfg.dt)
        arr.append([int(q["ID"]), int(q["IC"]), m["x"], m["y"]])
    it = iter(draws)
    used = 0
    ip = len(arr)
    while ip >= 1:
        pid, old_ic, x, y = arr[ip - 1]
        ic = int((x - cfg.xmin) / dx) + cfg.nx * int((y - cfg.ymin) / dy) + 1
        arr[ip - 1][1] = ic
        if ic != old_ic:
            ratio = cell_fnum[old_ic - 1] / cell_fnum[ic - 1]
            if ratio < 1.0:
                rfp = next(it); used += 1
                if ratio < rfp:
                    arr[ip - 1] = arr[-1]   # REMOVE_PARTICLE_ARRAY (src/particle.f90:139-179): the last one takes its place
                    arr.pop()
            elif ratio > 1.0:
                rfp = next(it); used += 1
                while (ratio - 1.0) > rfp:
                    arr.append([pid, ic, x, y])
                    ratio = ratio - 1.0
        ip -= 1
    return [a[0] for a in arr], [a[1] for a in arr], used


def test_radial_weighting_oracle_matches_fortran_reading(oracle):
    """BOOL_RADIAL_WEIGHTING, src/timecycle.f90:2110-2128, on scripted draws: same survivors, same clones, same array
    order (the reverse sweep and the swap-with-last removal), same number of draws."""
    rng = np.random.default_rng(12)
    nx, ny, n = 3, 4, 400
    w = wl.c1_argon_box(n, nx=nx, ny=ny, periodic=False)
    w.cfg.axi = 1
    w.cfg.ymin, w.cfg.ymax, w.cfg.zmin, w.cfg.zmax = 0.0, 1.0, 0.0, 1.0
    w.cfg.xmin, w.cfg.xmax = 0.0, 1.0
    w.cfg.bool_z_periodic = 0
    w.cfg.collision_type = capi.NO_COLL
    w.cfg.dt = 2.0e-4
    w.cfg.track_id = 1
    w.cfg.capacity = 8 * n
    p = w.particles
    p["X"] = rng.uniform(0.3, 0.7, n)           # (stay clear of the walls: the pin is about the weighting)
    p["Y"] = rng.uniform(0.3, 0.7, n)
    p["Z"] = 0.0
    p["VX"] = rng.normal(0.0, 300.0, n)
    p["VY"] = rng.normal(0.0, 300.0, n)
    p["VZ"] = rng.normal(0.0, 300.0, n)
    p["DTRIM"] = w.cfg.dt
    p["ID"] = np.arange(1, n + 1)
    dx, dy = 1.0 / nx, 1.0 / ny
    p["IC"] = (p["X"] / dx).astype(int) + nx * (p["Y"] / dy).astype(int) + 1
    row = np.array([0.5, 1.0, 3.7, 9.0])        # CELL_FNUM per cell row: ratios below 1, and above 2 (several clones)
    cell_fnum = np.repeat(row, nx)
    draws = rng.uniform(0.0, 1.0, 2 * n)
    ids, ics, used = radial_reading(p, cell_fnum, w.cfg, draws)
    wl.apply(oracle, w)
    oracle.set_cell_fnum(cell_fnum)
    script(oracle, draws)
    oracle.advect(1)
    out = oracle.download_particles()
    assert used > 50 and len(ids) != n
    assert script_used(oracle) == used
    # a clone is a new particle (INIT_PARTICLE draws a fresh ID, src/particle.f90:75-76) with its parent's state: VX is
    # untouched by the axisymmetric move and identifies the parent
    parent_of = {float(v): int(i) for v, i in zip(p["VX"], p["ID"])}
    parents = [parent_of[float(v)] for v in out["VX"]]
    assert parents == ids
    assert list(out["IC"]) == ics
    assert np.array_equal(out["ID"][out["ID"] <= n], np.array([i for k, i in enumerate(ids) if out["ID"][k] <= n]))
    assert max(np.bincount(parents)) >= 3       # some particle was cloned at least twice


# ------------------------------------------------------------------------------------------ Boris push
def boris_reading(v_old, charge, mass, e, b, dtime):
    """src/timecycle.f90:2271-2291 UPDATE_VELOCITY_BORIS, element by element in Fortran's evaluation order
    (left to right, no contraction); CROSS is src/tools.f90:1284-1292; QE from src/global.f90:35."""
    QE = 1.602176634e-19
    qc = charge * QE
    cross = lambda a, c: [a[1] * c[2] - a[2] * c[1], a[2] * c[0] - a[0] * c[2], a[0] * c[1] - a[1] * c[0]]
    v_minus = [v_old[k] + 0.5 * qc * e[k] / mass * dtime for k in range(3)]
    t = [0.5 * qc * b[k] / mass * dtime for k in range(3)]
    c = cross(v_minus, t)
    v_prime = [v_minus[k] + c[k] for k in range(3)]
    den = 1.0 + (t[0] * t[0] + t[1] * t[1] + t[2] * t[2])
    s = [2.0 * t[k] / den for k in range(3)]
    c = cross(v_prime, s)
    v_plus = [v_minus[k] + c[k] for k in range(3)]
    return [v_plus[k] + 0.5 * qc * e[k] / mass * dtime for k in range(3)]


def check_boris(lib):
    """One electron and one ion in the middle of a segment of a 1-D mesh, cell-constant E with three components
    (APPLY_E_FIELD, unstructured branch, src/fields.f90:4194-4201), uniform external B, a step short enough that no face
    is reached: V_NEW from the push, then X = X + V_NEW * DT (src/timecycle.f90:1229-1250, 2297-2310), bit for bit."""
    w = wl.c4_ccp_1d(ncells=4, n_per_species=1)
    w.cfg.collision_type = capi.NO_COLL
    w.cfg.dt = 3.0e-11
    bext = [0.013, -0.021, 0.034]
    for i, v in enumerate(bext):
        w.cfg.external_b[i] = v
    p = w.particles
    h = 0.0254 / 4
    p["X"] = [1.5 * h, 2.5 * h]
    p["Y"], p["Z"] = [0.41, 0.52], [0.63, 0.37]
    p["VX"], p["VY"], p["VZ"] = [8.1e5, -3.3e2], [-4.7e5, 2.9e2], [2.2e5, 1.7e2]
    p["S_ID"] = [w.species.id("e"), w.species.id("Ar+")]
    p["IC"] = [2, 3]
    p["DTRIM"] = w.cfg.dt
    ef = np.zeros((3, 4), order="F")
    ef[:, 1] = [1234.5, -987.6, 432.1]
    ef[:, 2] = [-2222.2, 333.3, 4444.4]
    w.efield = ef
    wl.apply(lib, w)
    lib.advect(1)
    out = lib.download_particles()
    assert len(out) == 2
    for k in range(2):
        s = int(p["S_ID"][k]) - 1
        vn = boris_reading([float(p["VX"][k]), float(p["VY"][k]), float(p["VZ"][k])], float(w.species.charge[s]),
                           float(w.species.mass[s]), [float(x) for x in ef[:, int(p["IC"][k]) - 1]], bext, w.cfg.dt)
        assert [float(out["VX"][k]), float(out["VY"][k]), float(out["VZ"][k])] == vn, (k, vn)
        assert float(out["X"][k]) == float(p["X"][k]) + vn[0] * w.cfg.dt
        assert float(out["Y"][k]) == float(p["Y"][k]) + vn[1] * w.cfg.dt
        assert float(out["Z"][k]) == float(p["Z"][k]) + vn[2] * w.cfg.dt
        assert out["IC"][k] == p["IC"][k]


def test_boris_push_oracle_matches_fortran_reading(oracle):
    check_boris(oracle)


@pytest.mark.gpu
def test_boris_push_device_matches_fortran_reading(gpu):
    check_boris(gpu)


# ------------------------------------------------------------------------------------------ DEPOSIT_CHARGE
def check_deposit_single_particle(lib):
    """src/fields.f90:5440-5520 for ONE charged particle (no summation order involved), in plain Python:
    triangle mesh (DIMS == 2): RHS(VP) = K*CHARGE*FNUM/(ZMAX-ZMIN) * (B1*X + B2*Y + B3) on the three nodes of its cell;
    uniform rectilinear grid: COMPUTE_WEIGHTS (:5341-5431) and RHS(INDICES(k)) = -K*CHARGE*FNUM/CELL_VOL * WEIGHTS(k).
    K = QE/(EPS0*EPS_SCALING**2), constants from src/global.f90:32-35."""
    QE, EPS0 = 1.602176634e-19, 8.8541878128e-12
    sp = wl.load_species("ar", ["Ar+", "e", "Ar"])
    # --- triangles
    w = wl.c3_tri_box(nx=4, ny=3, n=3, lo=(0.0, 0.0), hi=(0.1, 0.06), dt=1e-10, wall_bc=("specular",) * 4)
    w.species = sp
    w.cfg.pic_type = capi.PIC_EXPLICIT
    w.cfg.fnum = 3.5e4
    w.cfg.eps_scaling = 1.7
    m = w.mesh
    ic = 7
    nodes = np.asarray(m.cell_nodes)[:, ic - 1]
    xy = np.asarray(m.node_coords)[:2, nodes - 1]
    x, y = (float(v) for v in xy.mean(axis=1) + np.array([1.3e-3, -0.7e-3]))
    p = w.particles
    p["X"], p["Y"], p["Z"] = [x, x, x], [y, y, y], 0.0
    p["S_ID"] = [sp.id("Ar"), sp.id("e"), sp.id("Ar")]     # the neutrals are skipped (ABS(CHARGE) < 1e-6)
    p["IC"] = ic
    wl.apply(lib, w)
    rhs = lib.deposit_charge(m.nn)
    K = QE / (EPS0 * w.cfg.eps_scaling ** 2)
    rho_q = K * float(sp.charge[sp.id("e") - 1]) * w.cfg.fnum / (w.cfg.zmax - w.cfg.zmin)
    want = np.zeros(m.nn)
    B = np.asarray(m.basis_coeffs)
    for pp in range(3):
        psip = float(B[0, pp, ic - 1]) * x + float(B[1, pp, ic - 1]) * y + float(B[2, pp, ic - 1])
        want[nodes[pp] - 1] = 0.0 + rho_q * psip
    assert np.array_equal(rhs, want) and np.count_nonzero(rhs) == 3
    assert abs(rhs.sum() - rho_q) < 1e-12 * abs(rho_q)      # the P1 basis is a partition of unity
    # --- uniform rectilinear grid, DIMS == 2
    nx, ny = 5, 4
    cfg = capi.make_config(dims=2, nx=nx, ny=ny, xmin=-0.2, xmax=0.8, ymin=0.1, ymax=0.5, zmin=0.0, zmax=0.3, dt=1e-10,
                           fnum=7.7e3, pic_type=capi.PIC_EXPLICIT, bool_periodic=[1] * 4, bool_z_periodic=1)
    q = wl.new_particles(1)
    xp, yp = 0.3137, 0.2719
    hx, hy = (cfg.xmax - cfg.xmin) / float(nx), (cfg.ymax - cfg.ymin) / float(ny)
    ic = int((xp - cfg.xmin) / hx) + nx * int((yp - cfg.ymin) / hy) + 1
    q["X"], q["Y"], q["Z"], q["S_ID"], q["IC"], q["DTRIM"] = xp, yp, 0.1, sp.id("Ar+"), ic, cfg.dt
    w2 = wl.Workload("deposit-pin", cfg, sp, q)
    wl.apply(lib, w2)
    nn = (nx + 1) * (ny + 1)
    rhs = lib.deposit_charge(nn)
    i, j = (ic - 1) % nx, (ic - 1) // nx
    fx = (xp - (i * hx + cfg.xmin)) / hx
    fy = (yp - (j * hy + cfg.ymin)) / hy
    cell_vol = (cfg.xmax - cfg.xmin) * (cfg.ymax - cfg.ymin) * (cfg.zmax - cfg.zmin) / (nx * ny)
    rho_q = -(QE / (EPS0 * 1.0 ** 2)) * float(sp.charge[sp.id("Ar+") - 1]) * cfg.fnum / cell_vol
    i1 = ic + j - 1                                         # INDICES are 0-based node numbers (RHS(0:NNODES-1))
    want = np.zeros(nn)
    want[i1] = rho_q * ((1. - fx) * (1. - fy))
    want[i1 + 1] = rho_q * (fx * (1. - fy))
    want[i1 + nx + 1] = rho_q * ((1. - fx) * fy)
    want[i1 + nx + 2] = rho_q * (fx * fy)
    assert np.array_equal(rhs, want)


def test_deposit_single_particle_oracle_matches_fortran_reading(oracle):
    check_deposit_single_particle(oracle)


@pytest.mark.gpu
def test_deposit_single_particle_device_matches_fortran_reading(gpu):
    check_deposit_single_particle(gpu)


# ------------------------------------------------------------------------------------------ MCC_COLLISIONS_VAHEDI
def test_mcc_vahedi_oracle_matches_fortran_reading(oracle):
    """src/collisions.f90:1606-1892 on scripted draws: a three-product LXCat ionisation (partner materialised, energy
    redistribution with TOTDOF as the source counts it, two HS_SCATTERs, third product appended), a charge exchange, and
    a particle that fails the null-collision test; INTERP_CS (src/tools.f90:1544-1593) on an 8-point table."""
    t = PINS["mcc_vahedi"]
    sp = wl.load_species("ar", ["Ar", "Ar+", "e"])
    cfg = capi.make_config(dims=2, grid_type=capi.RECTILINEAR_UNIFORM, nx=1, ny=1, xmin=0.0, xmax=1.0, ymin=0.0, ymax=1.0,
                           zmin=-0.5, zmax=0.5, dt=fh(t["dt"]), fnum=1.0, collision_type=capi.MCC_VAHEDI,
                           mcc_bg_dens=fh(t["bg_dens"]), mcc_bg_ttra=fh(t["bg_ttra"]), bool_periodic=[1] * 4,
                           bool_z_periodic=1, capacity=64)
    p = wl.new_particles(len(t["inp"]))
    for k, q in enumerate(t["inp"]):
        p["X"][k], p["Y"][k], p["Z"][k] = q["pos"]
        p["VX"][k], p["VY"][k], p["VZ"][k] = (fh(c) for c in q["v"])
        p["EROT"][k], p["EVIB"][k] = fh(q["erot"]), fh(q["evib"])
        p["S_ID"][k] = sp.id(q["sid"])
    p["IC"] = 1
    p["DTRIM"] = cfg.dt
    reactions = [dict(type=capi.LXCAT, r1=sp.id(r["r1"]), r2=sp.id(r["r2"]), p1=sp.id(r["p1"]), p2=sp.id(r["p2"]),
                      p3=sp.id(r["p3"]) if r["p3"] else -1, p4=-1, nprod=r["nprod"], is_cex=r["is_cex"], ea=fh(r["ea"]),
                      table_energy=np.array([fh(v) for v in r["te"]]), table_cs=np.array([fh(v) for v in r["tcs"]]))
                 for r in t["reactions"]]
    w = wl.Workload("pin mcc", cfg, sp, p, reactions=reactions, mcc_background=([sp.id("Ar")], [1.0]))
    wl.apply(oracle, w)
    script(oracle, [fh(v) for v in t["draws"]])
    nc, nr = oracle.mcc_vahedi(1)
    assert (nc, nr) == (t["ncoll"], t["nreac"])
    assert script_used(oracle) == len(t["draws"])
    out = oracle.download_particles()
    assert len(out) == len(t["out"])
    for k, e in enumerate(t["out"]):
        assert out["S_ID"][k] == sp.id(e["sid"]), (k, out["S_ID"][k], e["sid"])
        got = [out["VX"][k], out["VY"][k], out["VZ"][k], out["EROT"][k], out["EVIB"][k]]
        want = [fh(c) for c in e["v"]] + [fh(e["erot"]), fh(e["evib"])]
        assert got == pytest.approx(want, rel=1e-13, abs=1e-300), (k, got, want)
        assert [out["X"][k], out["Y"][k], out["Z"][k]] == e["pos"] and out["IC"][k] == 1


# ------------------------------------------------------------------------------------------ BOUNDARIES_EMIT
def test_boundaries_emit_oracle_matches_fortran_reading(oracle):
    """src/timecycle.f90:368-547 (DIMS == 3) on scripted draws: fractional NFS per mixture component, MAXWELL tangentials
    (first component discarded), INTERNAL_ENERGY at TROT / TVIB, FLX with the drift ratio U_NORM*BETA, DTFRAC, the
    folded (S, T) point on the face; particles in creation order."""
    t = PINS["boundaries_emit"]
    w = tet_workload([], fh(t["dt"]))
    w.cfg.capacity = 64
    wl.apply(oracle, w)
    sp = w.species
    oracle.set_mixtures([[(sp.id(name), frac) for name, frac in t["mix"]]])
    oracle.set_species_weights([1.0] * len(sp.mass))
    oracle.set_emit_tasks([dict(t["task"], mix_id=0, iv1=1, iv2=1)])
    script(oracle, [fh(v) for v in t["draws"]])
    oracle.boundaries_emit(1)
    assert script_used(oracle) == t["used"]
    out = oracle.download_particles()
    assert len(out) == len(t["out"]) >= 3
    for k, e in enumerate(t["out"]):
        assert out["S_ID"][k] == sp.id(e["sid"]) and out["IC"][k] == e["ic"]
        for f in ("x", "y", "z", "vx", "vy", "vz", "erot", "evib", "dtrim"):
            assert float(out[f.upper()][k]) == pytest.approx(fh(e[f]), rel=1e-13, abs=1e-300), (k, f)


# ------------------------------------------------------------------------------------------ INITIAL_SEED
def test_initial_seed_oracle_matches_fortran_reading(oracle):
    """src/initialization.f90:2017-2233 (unstructured, DIMS == 3) on scripted draws: RANDINT per (task, component, cell),
    the folded (S, T, U) point in the tetrahedron, MAXWELL with a drift and a zero temperature component (no draw for it),
    INTERNAL_ENERGY, SPWT; particles in creation order."""
    t = PINS["initial_seed"]
    w = tet_workload([], 1.0e-6)
    w.species = wl.load_species("air", ["N2", "O2", "O"])
    w.cfg.fnum = fh(t["fnum"])
    w.cfg.capacity = 256
    wl.apply(oracle, w)
    sp = w.species
    oracle.set_mixtures([[(sp.id(name), frac) for name, frac in m] for m in t["mixtures"]])
    oracle.set_species_weights([t["spwt"][n] for n in ("N2", "O2", "O")])
    oracle.set_initial_particles_tasks([dict(q, rate=q["nrho"]) for q in t["tasks"]])
    script(oracle, [fh(v) for v in t["draws"]])
    oracle.initial_seed()
    assert script_used(oracle) == t["used"]
    out = oracle.download_particles()
    assert len(out) == len(t["out"])
    for k, e in enumerate(t["out"]):
        assert out["S_ID"][k] == sp.id(e["sid"]) and out["IC"][k] == e["ic"], k
        for f in ("x", "y", "z", "vx", "vy", "vz", "erot", "evib"):
            assert float(out[f.upper()][k]) == pytest.approx(fh(e[f]), rel=1e-13, abs=1e-300), (k, f)
        assert float(out["DTRIM"][k]) == 1.0e-6


# ------------------------------------------------------------------------------------------ VOLUME_INJECT
@pytest.mark.parametrize("case", [0, 1])
def test_volume_inject_structured_oracle_matches_fortran_reading(oracle, case):
    """src/timecycle.f90:553-737, structured branch on scripted draws; case 1 is axisymmetric (Y from the area-weighted
    law, Z = 0, DOMAIN_VOLUME with the 0.5*(YMAX**2-YMIN**2) factor)."""
    t = PINS["volume_inject"][case]
    g = t["grid"]
    sp = wl.load_species("air", ["N2", "O2", "O"])
    cfg = capi.make_config(dims=2, grid_type=capi.RECTILINEAR_UNIFORM, nx=g["nx"], ny=g["ny"], xmin=g["xmin"], xmax=g["xmax"],
                           ymin=g["ymin"], ymax=g["ymax"], zmin=g["zmin"], zmax=g["zmax"], dt=g["dt"], fnum=g["fnum"], axi=g["axi"],
                           collision_type=capi.NO_COLL, capacity=256)
    w = wl.Workload("pin vinj", cfg, sp, wl.new_particles(0))
    wl.apply(oracle, w)
    oracle.set_mixtures([[(sp.id(name), frac) for name, frac in m] for m in t["mixtures"]])
    oracle.set_species_weights([t["spwt"][n] for n in ("N2", "O2", "O")])
    oracle.set_volume_inject_tasks([dict(q, rate=q["nrhodot"]) for q in t["tasks"]])
    script(oracle, [fh(v) for v in t["draws"]])
    oracle.volume_inject(1)
    assert script_used(oracle) == t["used"]
    out = oracle.download_particles()
    assert len(out) == len(t["out"])
    for k, e in enumerate(t["out"]):
        assert out["S_ID"][k] == sp.id(e["sid"]) and out["IC"][k] == e["ic"], k
        for f in ("x", "y", "z", "vx", "vy", "vz", "erot", "evib"):
            assert float(out[f.upper()][k]) == pytest.approx(fh(e[f]), rel=1e-13, abs=1e-300), (k, f)
