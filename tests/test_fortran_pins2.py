"""Independent pins, second set: vectors evaluated in plain Python from the FORTRAN source (tests/golden/fortran_pins2.py
-> tests/golden/fortran_pins2.json), replayed through the oracle and, for the deterministic events, through the CUDA
library.  Reference lines covered: the header of tests/golden/fortran_pins2.py and DESIGN.md section 5.
"""
import json
import os
import types

import numpy as np
import pytest

from pantera_pic_dsmc_b200 import capi, workloads as wl

import util
from test_fortran_pins import script, script_used

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
PINS = json.load(open(os.path.join(GOLD, "fortran_pins2.json")))
fh = float.fromhex


def pin_tri_mesh():
    """The pin's own geometry (EDGE_NORMAL, CELL_EDGES_LEN evaluated by the pin from the Fortran formulas), not meshgen's."""
    m = PINS["tri_mesh"]
    nn, nc = len(m["nodes"]), len(m["cells"])
    o = types.SimpleNamespace(dims=2, nn=nn, nc=nc)
    xyz = np.zeros((3, nn))
    xyz[:2] = np.array(m["nodes"], dtype=np.float64).T
    o.node_coords = np.asfortranarray(xyz)
    o.cell_nodes = np.asfortranarray(np.array(m["cells"], dtype=np.int32).T)
    o.cell_neighbors = np.asfortranarray(np.array(m["neighbors"], dtype=np.int32).T)
    o.cell_faces_pg = np.asfortranarray(np.array(m["faces_pg"], dtype=np.int32).T)
    fnorm = np.zeros((3, 3, nc))
    fnorm[:2] = np.array(m["normal"], dtype=np.float64).transpose(2, 1, 0)
    o.face_normal = np.asfortranarray(fnorm)
    o.cell_edges_len = np.asfortranarray(np.array(m["edges_len"], dtype=np.float64).T)
    xy = np.array(m["nodes"])
    tri = np.array(m["cells"]) - 1
    a, b, c = xy[tri[:, 0]], xy[tri[:, 1]], xy[tri[:, 2]]
    area = 0.5 * np.abs(a[:, 0] * (b[:, 1] - c[:, 1]) + b[:, 0] * (c[:, 1] - a[:, 1]) + c[:, 0] * (a[:, 1] - b[:, 1]))
    o.cell_volumes = area * (a[:, 1] + b[:, 1] + c[:, 1]) / 3.0
    o.face_tang1 = o.face_tang2 = None
    o.cell_pg = None
    o.basis_coeffs = None
    o.cell_procs = None
    o.face_boundary_index = None
    o.nboundcells = 0
    o.boundary_area = None
    return o


def axi_workload():
    t = PINS["axi_walk"]
    sp = wl.load_species("air", ["N2", "O2"])
    dt = fh(t["dt"])
    cfg = capi.make_config(dims=2, axi=1, grid_type=capi.UNSTRUCTURED, nx=1, ny=1, xmin=0.0, xmax=0.09, ymin=0.01, ymax=0.05,
                           zmin=0.0, zmax=1.0, dt=dt, fnum=1.0, collision_type=capi.NO_COLL, dump_bound_avg_every=0)
    p = wl.new_particles(len(t["cases"]))
    for k, c in enumerate(t["cases"]):
        q = c["inp"]
        for f in ("x", "y", "z", "vx", "vy", "vz"):
            p[f.upper()][k] = fh(q[f])
        p["IC"][k] = q["ic"]
    p["S_ID"] = 1
    p["DTRIM"] = dt
    bcs = dict(particle_bc=[capi.SPECULAR] * 4, wall_temp=[300.0] * 4, acc_n=[1.0] * 4, acc_t=[1.0] * 4)
    w = wl.Workload("pin axisymmetric triangles", cfg, sp, p, mesh=pin_tri_mesh(), bcs=bcs)
    w.cfg.track_id = 1
    return w, t


def check_axi_walk(lib, exact):
    w, t = axi_workload()
    wl.apply(lib, w)
    lib.advect(1)
    out = util.by_id(lib.download_particles())
    assert len(out) == len(t["cases"])
    for k, c in enumerate(t["cases"]):
        e = c["out"]
        assert out["IC"][k] == e["ic"], (k, out["IC"][k], e["ic"], c["trace"])
        for f in ("x", "y", "z", "vx", "vy", "vz"):
            got, want = float(out[f.upper()][k]), fh(e[f])
            if exact:
                assert got == want, (k, f, got.hex(), e[f], c["trace"])
            else:
                assert got == pytest.approx(want, rel=1e-11, abs=1e-13), (k, f, got, want)


def test_axisymmetric_edge_walk_oracle_matches_fortran_reading(oracle):
    """Quadratic edge test, trial moves forth and back (their rounding stays in the particle), vertical edges, hops and
    specular reflections on all four walls: bit-identical with the plain-Python reading of src/timecycle.f90:1173-1253."""
    cases = PINS["axi_walk"]["cases"]
    assert max(len(c["trace"]) for c in cases) >= 7 and sum(c["trial_moves"] for c in cases) > 40
    check_axi_walk(oracle, exact=True)


@pytest.mark.gpu
def test_axisymmetric_edge_walk_device_matches_fortran_reading(gpu):
    """(the device evaluates the same expressions; the tolerance covers its fused trial move, see pgpu_advect.cu)"""
    check_axi_walk(gpu, exact=False)


def test_dsmc_collisions_visit_cells_and_particles_in_the_reference_order(oracle):
    """Two cells, store order interleaved: the scripted INT(NPC*rf()) picks must land on the particles the stable
    counting sort of DSMC_COLLISIONS puts at those places (src/collisions.f90:37-149, 259-271)."""
    t = PINS["dsmc_order"]
    w = wl.c1_argon_box(7, nx=2, ny=1)
    w.cfg.fnum, w.cfg.dt = fh(t["fnum"]), fh(t["dt"])
    assert fh(t["sigmamax"]) == pytest.approx(w.species.sigmamax, rel=1e-14)
    p = wl.new_particles(7)
    for k, (cell, v) in enumerate(zip(t["cells"], t["inp"])):
        p["X"][k] = 0.25 + 0.5 * (cell - 1)
        p["Y"][k] = 0.1 + 0.1 * k
        p["VX"][k], p["VY"][k], p["VZ"][k] = (fh(c) for c in v)
        p["IC"][k] = cell
    p["S_ID"] = 1
    p["DTRIM"] = w.cfg.dt
    w.particles = p
    w.cfg.track_id = 1
    wl.apply(oracle, w)
    script(oracle, t["draws"])
    ncoll, nreac = oracle.collide(1)
    assert (ncoll, nreac) == (2, 0) and script_used(oracle) == len(t["draws"])
    out = util.by_id(oracle.download_particles())
    for k in range(7):
        for f, want in zip(("VX", "VY", "VZ"), t["out"][k]):
            got = float(out[f][k])
            if k in t["changed"]:
                assert got == pytest.approx(fh(want), rel=1e-13), (k, f)
            else:
                assert got == fh(want), (k, f)


def check_cell_nonuniform(lib):
    """CELL_FROM_POSITION on a non-uniform rectilinear grid is what re-indexes a particle after its flight
    (src/timecycle.f90:2104-2106): particles at rest are given a wrong IC and one ADVECT must bring the pinned one back."""
    t = PINS["cell_nonuniform"]
    xc, yc = t["xcoord"], t["ycoord"]
    nx, ny = len(xc) - 1, len(yc) - 1
    w = wl.c1_argon_box(len(t["cases"]), nx=nx, ny=ny)
    w.cfg.grid_type = capi.RECTILINEAR_NONUNIFORM
    w.cfg.xmin, w.cfg.xmax, w.cfg.ymin, w.cfg.ymax = xc[0], xc[-1], yc[0], yc[-1]
    w.cfg.collision_type = capi.NO_COLL
    w.cfg.track_id = 1
    p = wl.new_particles(len(t["cases"]))
    for k, c in enumerate(t["cases"]):
        p["X"][k], p["Y"][k] = fh(c["x"]), fh(c["y"])
        p["IC"][k] = 1 + (k * 7) % (nx * ny)
    p["S_ID"] = 1
    p["DTRIM"] = w.cfg.dt
    w.particles = p
    wl.apply(lib, w, upload=False)
    lib.set_grid_coords(xc, yc)
    lib.upload_particles(p)
    lib.advect(1)
    out = util.by_id(lib.download_particles())
    assert [int(i) for i in out["IC"]] == [c["ic"] for c in t["cases"]]
    assert np.array_equal(out["X"], p["X"]) and np.array_equal(out["Y"], p["Y"])


def test_cell_from_position_nonuniform_oracle_matches_fortran_reading(oracle):
    check_cell_nonuniform(oracle)


@pytest.mark.gpu
def test_cell_from_position_nonuniform_device_matches_fortran_reading(gpu):
    check_cell_nonuniform(gpu)


# ---- rectilinear grid: box sides, interior walls, Z wrap, re-index -------------------------------------------------
def structured_workload(cases):
    t = PINS["structured_walk"]
    g = t["grid"]
    sp = wl.load_species("air", ["N2", "O2"])
    dt = fh(t["dt"])
    cfg = capi.make_config(dims=2, grid_type=capi.RECTILINEAR_UNIFORM, nx=g["nx"], ny=g["ny"], xmin=g["xmin"], xmax=g["xmax"],
                           ymin=g["ymin"], ymax=g["ymax"], zmin=g["zmin"], zmax=g["zmax"], dt=dt, fnum=1.0, collision_type=capi.NO_COLL,
                           bool_periodic=g["periodic"], bool_specular=g["specular"], bool_diffuse=g["diffuse"], boundtemp=g["boundtemp"],
                           bool_z_periodic=g["z_periodic"])
    p = wl.new_particles(len(cases))
    for k, c in enumerate(cases):
        q = c["inp"]
        for f in ("x", "y", "z", "vx", "vy", "vz", "erot", "evib"):
            p[f.upper()][k] = fh(q[f])
        p["S_ID"][k] = sp.id(q["sid"])
        ddx, ddy = (g["xmax"] - g["xmin"]) / g["nx"], (g["ymax"] - g["ymin"]) / g["ny"]
        p["IC"][k] = int((fh(q["x"]) - g["xmin"]) / ddx) + g["nx"] * int((fh(q["y"]) - g["ymin"]) / ddy) + 1
    p["DTRIM"] = dt
    ws = t["walls"]
    col = lambda k: [w[k] for w in ws]
    walls = dict(x1=col("x1"), y1=col("y1"), x2=col("x2"), y2=col("y2"), normx=col("normx"), normy=col("normy"), temp=col("temp"),
                 specular=col("specular"), diffuse=col("diffuse"), porous=col("porous"), transmissivity=col("transm"), react=[0] * len(ws))
    w = wl.Workload("pin rectilinear grid", cfg, sp, p, walls=walls)
    w.cfg.track_id = 1
    return w


def check_structured(lib, cases, exact, rel=1e-13):
    out = util.by_id(lib.download_particles())
    alive = [c for c in cases if c["out"] is not None]
    assert len(out) == len(alive)
    for k, c in enumerate(alive):
        e = c["out"]
        assert out["IC"][k] == e["ic"], (k, out["IC"][k], e["ic"], c["events"])
        for f in ("x", "y", "z", "vx", "vy", "vz", "erot", "evib"):
            got, want = float(out[f.upper()][k]), fh(e[f])
            if exact:
                assert got == want, (k, f, got.hex(), e[f], c["events"])
            else:
                assert got == pytest.approx(want, rel=rel, abs=1e-300), (k, f, got, want, c["events"])


def check_structured_det(lib):
    cases = PINS["structured_walk"]["det"]
    w = structured_workload(cases)
    wl.apply(lib, w)
    lib.advect(1)
    check_structured(lib, cases, exact=True)
    b, wc = lib.get_boundary_counts(2, nw=3)
    assert list(b[:4]) == [1, 1, 2, 0] and list(wc[:3]) == [1, 0, 0]      # LOCAL_BOUNDARY_COLL_COUNT / LOCAL_WALL_COLL_COUNT of N2


def test_structured_walk_oracle_matches_fortran_reading(oracle):
    """Periodic and specular box sides (strict '>'), a one-sided interior wall with the hit point inside / outside the
    segment, the Z wrap and the final CELL_FROM_POSITION: bit-identical."""
    check_structured_det(oracle)


@pytest.mark.gpu
def test_structured_walk_device_matches_fortran_reading(gpu):
    check_structured_det(gpu)


def test_structured_walls_with_draws_oracle_matches_fortran_reading(oracle):
    """Porous specular wall (removed / reflected by the scripted draw), diffuse box side at BOUNDTEMP, diffuse interior wall."""
    t = PINS["structured_walk"]["sto"]
    w = structured_workload(t["cases"])
    wl.apply(oracle, w)
    script(oracle, t["draws"])
    oracle.advect(1)
    assert script_used(oracle) == t["used"] == len(t["draws"])
    # the removed particle is the first of the store: the swap-with-last removal moves the last one into its place;
    # the IDs keep the match
    out = util.by_id(oracle.download_particles())
    alive = [c for c in t["cases"] if c["out"] is not None]
    assert len(out) == 3 and len(alive) == 3
    check_structured(oracle, t["cases"], exact=False, rel=1e-13)


# ---- MCC_COLLISIONS ------------------------------------------------------------------------------------------------
def test_mcc_collisions_oracle_matches_fortran_reading(oracle):
    """Three test particles against an O2/N2 background on a scripted stream: partner sampling before the test, partner and
    product creation order, the TCE dissociation N2 + O2 --> N2 + O + O of the fast particle, Borgnakke-Larsen and VSS
    deflection of the others (src/collisions.f90:1261-1595)."""
    t = PINS["mcc_collisions"]
    w = wl.mcc_box("air", n=3, nx=1, ny=1, bg_T=fh(t["bg_ttra"]), bg_dens=fh(t["bg_dens"]), dt=fh(t["dt"]))
    sp = w.species
    assert [sp.names[k] for k in range(5)] == t["species"]
    w.mcc_background = ([sp.id(n) for n, _ in t["bg_mix"]], [f for _, f in t["bg_mix"]])
    p = wl.new_particles(3)
    for k, q in enumerate(t["inp"]):
        p["X"][k], p["Y"][k], p["Z"][k] = q["pos"]
        p["VX"][k], p["VY"][k], p["VZ"][k] = (fh(c) for c in q["v"])
        p["EROT"][k], p["EVIB"][k] = fh(q["erot"]), fh(q["evib"])
        p["S_ID"][k] = sp.id(q["sid"])
    p["IC"] = 1
    p["DTRIM"] = w.cfg.dt
    w.particles = p
    wl.apply(oracle, w)
    script(oracle, [fh(v) for v in t["draws"]])
    ncoll, nreac = oracle.mcc_collisions(1)
    assert (ncoll, nreac) == (t["ncoll"], t["nreac"]) and script_used(oracle) == t["used"]
    out = oracle.download_particles()                       # store order: the 3 test particles, then partners / products as created
    assert len(out) == len(t["out"])
    for k, e in enumerate(t["out"]):
        assert out["S_ID"][k] == sp.id(e["sid"]), (k, e["sid"])
        for f, want in zip(("VX", "VY", "VZ"), e["v"]):
            assert float(out[f][k]) == pytest.approx(fh(want), rel=1e-12, abs=1e-9), (k, f)
        assert float(out["EROT"][k]) == pytest.approx(fh(e["erot"]), rel=1e-11, abs=1e-33), k
        assert float(out["EVIB"][k]) == pytest.approx(fh(e["evib"]), rel=1e-11, abs=1e-33), k
        assert (out["X"][k], out["Y"][k], out["Z"][k]) == tuple(e["pos"])


# ---- EXCHANGE --------------------------------------------------------------------------------------------------------
def fortran_exchange(stores, owner_of):
    """Plain-Python reading of EXCHANGE + COUNTING_SORT (src/grid_and_partition.f90:482-518, 525-622) for all ranks at once.
    stores[r]: list of (id, ic) in store order; owner_of(ic) = PROC_FROM_CELL.  Returns the stores after the exchange."""
    nr = len(stores)
    sendbufs, sendcounts = [], []
    for r in range(nr):
        p = list(stores[r])
        np_proc = len(p)
        sendbuf, sendcount = [], [0] * nr
        ip = np_proc
        while ip >= 1:                                   # from the last particle to the first
            jproc = owner_of(p[ip - 1][1])
            if jproc != r:
                sendcount[jproc] += 1
                sendbuf.append(p[ip - 1])
                p[ip - 1] = p[np_proc - 1]               # REMOVE_PARTICLE_ARRAY: the last one takes the place
                np_proc -= 1
            ip -= 1
        stores[r] = p[:np_proc]
        disp = list(sendcount)                           # COUNTING_SORT: inclusive displacements, filled from the back
        for i in range(1, nr):
            disp[i] = disp[i] + disp[i - 1]
        srt = [None] * len(sendbuf)
        for i in range(len(sendbuf), 0, -1):
            iproc = owner_of(sendbuf[i - 1][1])
            srt[disp[iproc] - 1] = sendbuf[i - 1]
            disp[iproc] -= 1
        sendbufs.append(srt)
        sendcounts.append(sendcount)
    for r in range(nr):                                  # MPI_ALLTOALLV: blocks in source-rank order, then ADD_PARTICLE_ARRAY
        for s in range(nr):
            off = sum(sendcounts[s][:r])
            stores[r] = stores[r] + sendbufs[s][off:off + sendcounts[s][r]]
    return stores


def test_exchange_store_order_oracle_matches_fortran_reading():
    """The ORDER of a rank's particle array after EXCHANGE decides which pairs its sequential MT stream selects afterwards:
    reverse sweep with swap-with-last removal, stable counting sort of the send buffer, blocks appended in source-rank order."""
    from oracle.binding import load_oracle
    nr = 3
    rng = np.random.default_rng(31)
    libs, stores = [], []
    w0 = wl.c5_channel(6, 2, 2, ppc=1)
    mesh = w0.mesh
    from pantera_pic_dsmc_b200 import meshgen as mg
    cell_procs = mg.assign_cells_to_procs(mesh, nr, 0)
    assert set(np.unique(cell_procs)) == {0, 1, 2}
    for r in range(nr):
        w = wl.c5_channel(6, 2, 2, ppc=1)
        w.mesh.cell_procs = cell_procs
        w.cfg.rank, w.cfg.nranks, w.cfg.track_id = r, nr, 1
        n = 40 + 7 * r
        p = wl.new_particles(n)
        p["IC"] = rng.integers(1, mesh.nc + 1, n)         # anywhere in the mesh: about two thirds must be shipped
        p["S_ID"] = 1
        p["ID"] = 1000 * (r + 1) + np.arange(n)
        p["X"] = rng.random(n)
        w.particles = p
        o = load_oracle(rng_mode=0)
        wl.apply(o, w)
        libs.append(o)
        stores.append([(int(i), int(c)) for i, c in zip(p["ID"], p["IC"])])
    packed = [o.exchange_pack(nr) for o in libs]          # local half on every rank, then the all-to-all by hand
    for r in range(nr):
        for s in range(nr):
            if s == r:
                continue
            buf, cnt = packed[s]
            off = int(cnt[:r].sum())
            if cnt[r]:
                libs[r].append_particles(buf[off:off + int(cnt[r])])
    expect = fortran_exchange(stores, lambda ic: int(cell_procs[ic - 1]))
    moved = 0
    for r in range(nr):
        got = libs[r].download_particles()
        assert [int(i) for i in got["ID"]] == [i for i, _ in expect[r]], r
        assert all(cell_procs[ic - 1] == r for ic in got["IC"])
        moved += sum(1 for i, _ in expect[r] if i // 1000 != r + 1)
        libs[r].close()
    assert moved > 50


# ---- field gather + Boris kick -----------------------------------------------------------------------------------------
def cic_e_reading(cfg, ef, x, y, ic):
    """APPLY_E_FIELD on a uniform rectilinear 2-D grid (src/fields.f90:4209-4216) with COMPUTE_WEIGHTS (:5341-5404):
    E_FIELD(3, 0:NPY-1, 0:NPX-1) indexed (:, INDJ, INDI); ef is that array, shape (3, NPY, NPX)."""
    nx, ny = cfg.nx, cfg.ny
    hx = (cfg.xmax - cfg.xmin) / float(nx)
    hy = (cfg.ymax - cfg.ymin) / float(ny)
    i = (ic - 1) % nx
    j = (ic - 1) // nx
    cell_xmin = i * hx + cfg.xmin
    cell_ymin = j * hy + cfg.ymin
    fx = (x - cell_xmin) / hx
    fy = (y - cell_ymin) / hy
    wts = [(1.0 - fx) * (1.0 - fy), fx * (1.0 - fy), (1.0 - fx) * fy, fx * fy]
    indi = [i, i + 1, i, i + 1]
    indj = [j, j, j + 1, j + 1]
    e = []
    for c in range(3):
        e.append(wts[0] * ef[c, indj[0], indi[0]] + wts[1] * ef[c, indj[1], indi[1]] + wts[2] * ef[c, indj[2], indi[2]]
                 + wts[3] * ef[c, indj[3], indi[3]])
    return [float(v) for v in e]


def check_cic_gather(lib):
    """Ions and electrons on a 5 x 4 periodic grid, a nodal E with a different value in every node and component, a uniform
    external B, a step short enough that no side is reached: V_NEW = Boris(E gathered at the OLD position), X = X + V_NEW DT."""
    from test_fortran_pins import boris_reading
    sp = wl.load_species("ar", ["Ar+", "e"])
    nx, ny = 5, 4
    bext = [0.011, -0.007, 0.023]
    cfg = capi.make_config(dims=2, nx=nx, ny=ny, xmin=-0.2, xmax=0.8, ymin=0.1, ymax=0.5, dt=4e-11, pic_type=capi.PIC_EXPLICIT,
                           bool_periodic=[1] * 4, bool_z_periodic=1, external_b=bext, collision_type=capi.NO_COLL)
    rng = np.random.default_rng(77)
    n = 24
    p = wl.new_particles(n)
    p["X"] = cfg.xmin + (cfg.xmax - cfg.xmin) * (0.02 + 0.96 * rng.random(n))
    p["Y"] = cfg.ymin + (cfg.ymax - cfg.ymin) * (0.02 + 0.96 * rng.random(n))
    p["Z"] = 0.3 * (rng.random(n) - 0.5)
    p["S_ID"] = rng.integers(1, 3, n)
    p["VX"], p["VY"], p["VZ"] = (rng.standard_normal(n) * 3e4 for _ in range(3))
    p["DTRIM"] = cfg.dt
    hx, hy = (cfg.xmax - cfg.xmin) / nx, (cfg.ymax - cfg.ymin) / ny
    p["IC"] = ((p["X"] - cfg.xmin) / hx).astype(np.int32) + nx * ((p["Y"] - cfg.ymin) / hy).astype(np.int32) + 1
    ef = np.asfortranarray(rng.standard_normal((3, ny + 1, nx + 1)) * 2e4)
    w = wl.Workload("pin CIC gather", cfg, sp, p, efield=ef)
    w.cfg.track_id = 1
    wl.apply(lib, w)
    lib.advect(1)
    out = util.by_id(lib.download_particles())
    assert len(out) == n
    for k in range(n):
        s = int(p["S_ID"][k]) - 1
        e = cic_e_reading(cfg, ef, float(p["X"][k]), float(p["Y"][k]), int(p["IC"][k]))
        vn = boris_reading([float(p["VX"][k]), float(p["VY"][k]), float(p["VZ"][k])], float(sp.charge[s]), float(sp.mass[s]), e, bext, cfg.dt)
        assert [float(out["VX"][k]), float(out["VY"][k]), float(out["VZ"][k])] == vn, (k, vn)
        assert float(out["X"][k]) == float(p["X"][k]) + vn[0] * cfg.dt and float(out["Y"][k]) == float(p["Y"][k]) + vn[1] * cfg.dt


def test_cic_field_gather_oracle_matches_fortran_reading(oracle):
    check_cic_gather(oracle)


@pytest.mark.gpu
def test_cic_field_gather_device_matches_fortran_reading(gpu):
    check_cic_gather(gpu)


def check_nodal_b_gather(lib):
    """APPLY_B_FIELD (src/fields.f90:4227-4249): B = sum_j B_FIELD(:, 1, V_j) PSI_j(x, y) on the pin's triangles (not
    axisymmetric here), + EXTERNAL_B_FIELD, with the cell-constant E of the unstructured branch of APPLY_E_FIELD."""
    from test_fortran_pins import boris_reading
    m = pin_tri_mesh()
    mm = PINS["tri_mesh"]
    xy = np.array(mm["nodes"])
    tri = np.array(mm["cells"]) - 1
    # BASIS_COEFFS of the triangles, src/grid_and_partition.f90:1455-1476
    bc = np.zeros((3, 3, m.nc))
    for c in range(m.nc):
        (x1, y1), (x2, y2), (x3, y3) = xy[tri[c, 0]], xy[tri[c, 1]], xy[tri[c, 2]]
        area = 0.5 * abs(x1 * (y2 - y3) + x2 * (y3 - y1) + x3 * (y1 - y2))
        raw = [[y2 - y3, -(x2 - x3), x2 * y3 - x3 * y2], [-(y1 - y3), x1 - x3, x3 * y1 - x1 * y3], [-(y2 - y1), x2 - x1, x1 * y2 - x2 * y1]]
        for i in range(3):
            for j in range(3):
                bc[j, i, c] = 0.5 * raw[i][j] / area
    m.basis_coeffs = np.asfortranarray(bc)
    m.cell_volumes = np.array([0.5 * abs(xy[t[0], 0] * (xy[t[1], 1] - xy[t[2], 1]) + xy[t[1], 0] * (xy[t[2], 1] - xy[t[0], 1])
                                         + xy[t[2], 0] * (xy[t[0], 1] - xy[t[1], 1])) for t in tri])
    sp = wl.load_species("ar", ["Ar+", "e"])
    bext = [0.002, 0.004, -0.003]
    cfg = capi.make_config(dims=2, grid_type=capi.UNSTRUCTURED, nx=1, ny=1, xmin=0.0, xmax=0.09, ymin=0.01, ymax=0.05, zmin=-0.5, zmax=0.5,
                           dt=2e-11, fnum=1.0, collision_type=capi.NO_COLL, pic_type=capi.PIC_EXPLICIT, bool_z_periodic=1, external_b=bext,
                           dump_bound_avg_every=0)
    rng = np.random.default_rng(5)
    n = m.nc
    p = wl.new_particles(n)
    lam = rng.dirichlet([2.0, 2.0, 2.0], n)                    # one particle well inside every triangle
    p["X"] = (lam * xy[tri, 0]).sum(axis=1)
    p["Y"] = (lam * xy[tri, 1]).sum(axis=1)
    p["IC"] = np.arange(1, n + 1)
    p["S_ID"] = rng.integers(1, 3, n)
    p["VX"], p["VY"], p["VZ"] = (rng.standard_normal(n) * 2e4 for _ in range(3))
    p["DTRIM"] = cfg.dt
    bcs = dict(particle_bc=[capi.SPECULAR] * 4, wall_temp=[300.0] * 4, acc_n=[1.0] * 4, acc_t=[1.0] * 4)
    ef = np.asfortranarray(rng.standard_normal((3, m.nc)) * 1e4)
    bf = np.asfortranarray(rng.standard_normal((3, m.nn)) * 0.02)
    w = wl.Workload("pin nodal B", cfg, sp, p, mesh=m, bcs=bcs, efield=ef)
    w.cfg.track_id = 1
    wl.apply(lib, w)
    lib.set_bfield(bf)
    lib.advect(1)
    out = util.by_id(lib.download_particles())
    assert len(out) == n
    for k in range(n):
        c = k
        x, y = float(p["X"][k]), float(p["Y"][k])
        b = [0.0, 0.0, 0.0]
        for j in range(3):
            psij = x * bc[0, j, c] + y * bc[1, j, c] + bc[2, j, c]
            for a in range(3):
                b[a] = b[a] + bf[a, tri[c, j]] * psij
        b = [float(b[a] + bext[a]) for a in range(3)]
        s = int(p["S_ID"][k]) - 1
        vn = boris_reading([float(p["VX"][k]), float(p["VY"][k]), float(p["VZ"][k])], float(sp.charge[s]), float(sp.mass[s]),
                           [float(v) for v in ef[:, c]], b, cfg.dt)
        assert [float(out["VX"][k]), float(out["VY"][k]), float(out["VZ"][k])] == vn, (k, vn)
        assert out["IC"][k] == k + 1


def test_nodal_b_gather_oracle_matches_fortran_reading(oracle):
    check_nodal_b_gather(oracle)


@pytest.mark.gpu
def test_nodal_b_gather_device_matches_fortran_reading(gpu):
    check_nodal_b_gather(gpu)
