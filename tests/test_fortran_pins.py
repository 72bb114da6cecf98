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
