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


# ---- TALLY_PARTICLE_TO_BOUNDARY ------------------------------------------------------------------------------------------
def test_boundary_tallies_oracle_match_fortran_reading(oracle):
    """The tet walk of the first pin set once more, this time accumulating TIMESTEP_NIN/NOUT/P*IN/P*OUT/EIN/EOUT as
    TALLY_PARTICLE_TO_BOUNDARY does (src/postprocess.f90:1098-1147): the incident tally with the velocity before the wall
    model, the reflected one after it (src/timecycle.f90:1373-1376, 1616-1621); K = FNUM*MOLMASS/AREA/DT, signs, the
    (boundary element, species) index.  The walk below is a plain-Python reading of src/timecycle.f90:1257-1312, 1433-1443."""
    import test_fortran_pins as t1
    pins1 = t1.PINS
    mesh = pins1["tet_mesh"]
    cases = pins1["tet_walk"]["cases"]
    dt = fh(pins1["tet_walk"]["dt"])
    w = t1.tet_workload(cases, dt)
    w.cfg.fnum = 3.0e12
    w.cfg.dump_bound_start, w.cfg.dump_bound_avg_every = 0, 1
    w.cfg.track_id = 1
    nodes, cells = np.array(mesh["nodes"]), np.array(mesh["cells"])
    ind = [[1, 3, 2], [1, 2, 4], [2, 3, 4], [1, 4, 3]]
    bidx = np.full((4, len(cells)), -1, dtype=np.int32)
    areas = []
    for c in range(len(cells)):
        for j in range(4):
            if mesh["neighbors"][c][j] == -1:
                a, b, cc = (nodes[cells[c][ind[j][m] - 1] - 1] for m in range(3))
                cr = np.cross(b - a, cc - a)
                areas.append(0.5 * float(np.sqrt(cr[0] * cr[0] + cr[1] * cr[1] + cr[2] * cr[2])))
                bidx[j, c] = len(areas)
    nb = len(areas)
    w.mesh.face_boundary_index = np.asfortranarray(bidx)
    w.mesh.nboundcells = nb
    w.mesh.boundary_area = np.array(areas)
    wl.apply(oracle, w)
    oracle.advect(1)
    got = oracle.get_boundary_tallies(nb, 2)          # (10, ns, nboundcells): NIN, NOUT, PXIN, PYIN, PZIN, PXOUT, PYOUT, PZOUT, EIN, EOUT
    sp = w.species
    exp = np.zeros((10, 2, nb))
    hits = 0
    for c in cases:                                   # particles in store order
        q = {k: fh(v) if isinstance(v, str) and k != "sid" else v for k, v in c["inp"].items()}
        x, y, z, vx, vy, vz, ic = q["x"], q["y"], q["z"], q["vx"], q["vy"], q["vz"], q["ic"]
        js = sp.id(q["sid"]) - 1
        molmass = float(sp.mass[js])
        dtrim = dt
        while dtrim > 0.0:
            dtcoll, boundcoll = dtrim, -1
            for i in range(1, 5):
                n = mesh["normal"][ic - 1][i - 1]
                P = mesh["nodes"][mesh["cells"][ic - 1][i - 1] - 1]
                vn = vx * n[0] + vy * n[1] + vz * n[2]
                dx = (P[0] - x) * n[0] + (P[1] - y) * n[1] + (P[2] - z) * n[2]
                if vn >= 0.0 and vn * dtcoll >= dx:
                    dtcoll, boundcoll = dx / vn, i
            if boundcoll == -1:
                x, y, z = x + vx * dtrim, y + vy * dtrim, z + vz * dtrim
                dtrim = 0.0
                continue
            x, y, z = x + vx * dtcoll, y + vy * dtcoll, z + vz * dtcoll
            dtrim = dtrim - dtcoll
            nbr = mesh["neighbors"][ic - 1][boundcoll - 1]
            if nbr != -1:
                ic = nbr
                continue
            hits += 1
            b = bidx[boundcoll - 1, ic - 1] - 1
            area = areas[b]
            k = w.cfg.fnum * molmass / area / dt
            exp[0, js, b] = exp[0, js, b] + w.cfg.fnum / area / dt
            exp[2, js, b] = exp[2, js, b] + k * vx
            exp[3, js, b] = exp[3, js, b] + k * vy
            exp[4, js, b] = exp[4, js, b] + k * vz
            exp[8, js, b] = exp[8, js, b] + k * (vx * vx + vy * vy + vz * vz)
            fn = [-v for v in mesh["normal"][ic - 1][boundcoll - 1]]
            vdotn = vx * fn[0] + vy * fn[1] + vz * fn[2]
            vx, vy, vz = vx - 2.0 * vdotn * fn[0], vy - 2.0 * vdotn * fn[1], vz - 2.0 * vdotn * fn[2]
            exp[1, js, b] = exp[1, js, b] - w.cfg.fnum / area / dt
            exp[5, js, b] = exp[5, js, b] - k * vx
            exp[6, js, b] = exp[6, js, b] - k * vy
            exp[7, js, b] = exp[7, js, b] - k * vz
            exp[9, js, b] = exp[9, js, b] - k * (vx * vx + vy * vy + vz * vz)
    assert hits >= 5 and np.count_nonzero(exp[0]) >= 3
    assert np.array_equal(got, exp), np.abs(got - exp).max()


# ---- BASIS_COEFFS and DEPOSIT_CHARGE on tetrahedra and segments --------------------------------------------------------------
def basis_tet_reading(P, volume):
    """BASIS_COEFFS(4, 4) of one tetrahedron, src/grid_and_partition.f90:2013-2036 (P: 4 nodes x 3 coordinates)."""
    (X1, Y1, Z1), (X2, Y2, Z2), (X3, Y3, Z3), (X4, Y4, Z4) = ([float(v) for v in row] for row in P)
    B = [[0.0] * 4 for _ in range(4)]                      # B[j][i] = BASIS_COEFFS(j+1, i+1)
    B[0][0] = Y2*Z3-Y3*Z2 -Y2*Z4+Y4*Z2 +Y3*Z4-Y4*Z3
    B[1][0] = -X2*Z3+X3*Z2 +X2*Z4-X4*Z2 -X3*Z4+X4*Z3
    B[2][0] = X2*Y3-X3*Y2 -X2*Y4+X4*Y2 +X3*Y4-X4*Y3
    B[3][0] = -X2*Y3*Z4 +X3*Y2*Z4 +X2*Y4*Z3 -X4*Y2*Z3 -X3*Y4*Z2 +X4*Y3*Z2
    B[0][1] = -Y1*Z3+Y3*Z1 +Y1*Z4-Y4*Z1 -Y3*Z4+Y4*Z3
    B[1][1] = X1*Z3-X3*Z1 -X1*Z4+X4*Z1 +X3*Z4-X4*Z3
    B[2][1] = -X1*Y3+X3*Y1 +X1*Y4-X4*Y1 -X3*Y4+X4*Y3
    B[3][1] = X1*Y3*Z4 -X3*Y1*Z4 -X1*Y4*Z3 +X4*Y1*Z3 +X3*Y4*Z1 -X4*Y3*Z1
    B[0][2] = Y1*Z2-Y2*Z1 -Y1*Z4+Y4*Z1 +Y2*Z4-Y4*Z2
    B[1][2] = -X1*Z2+X2*Z1 +X1*Z4-X4*Z1 -X2*Z4+X4*Z2
    B[2][2] = X1*Y2-X2*Y1 -X1*Y4+X4*Y1 +X2*Y4-X4*Y2
    B[3][2] = -X1*Y2*Z4 +X2*Y1*Z4 +X1*Y4*Z2 -X4*Y1*Z2 -X2*Y4*Z1 +X4*Y2*Z1
    B[0][3] = -Y1*Z2+Y2*Z1 +Y1*Z3-Y3*Z1 -Y2*Z3+Y3*Z2
    B[1][3] = X1*Z2-X2*Z1 -X1*Z3+X3*Z1 +X2*Z3-X3*Z2
    B[2][3] = -X1*Y2+X2*Y1 +X1*Y3-X3*Y1 -X2*Y3+X3*Y2
    B[3][3] = X1*Y2*Z3 -X2*Y1*Z3 -X1*Y3*Z2 +X3*Y1*Z2 +X2*Y3*Z1 -X3*Y2*Z1
    return [[-B[j][i] / 6.0 / volume for i in range(4)] for j in range(4)]


def test_basis_coeffs_and_deposit_on_tets_and_segments_match_fortran_reading(oracle):
    """(1) meshgen's BASIS_COEFFS / CELL_VOLUMES (what READ_3D/1D_UNSTRUCTURED_GRID_SU2 would hand over) against the
    Fortran expressions evaluated here, on a jittered tet box and a line; the basis is the barycentric one (PSI_i(x_j) =
    delta_ij).  (2) DEPOSIT_CHARGE (src/fields.f90:5459-5487) of single charged particles on tetrahedra (RHO_Q =
    K*CHARGE*FNUM) and on segments (RHO_Q = K*CHARGE*FNUM/(YMAX-YMIN)/(ZMAX-ZMIN)), bit for bit."""
    QE, EPS0 = 1.602176634e-19, 8.8541878128e-12
    sp = wl.load_species("ar", ["Ar+", "e", "Ar"])
    # ---- tetrahedra
    w = wl.c5_channel(3, 2, 2, ppc=1, jitter=0.2)
    m = w.mesh
    xyz = np.asarray(m.node_coords)
    cn = np.asarray(m.cell_nodes) - 1
    B = np.asarray(m.basis_coeffs)
    for c in (0, 5, 17, m.nc - 1):
        P = xyz[:, cn[:, c]].T
        a, b, cc = P[1] - P[0], P[2] - P[0], P[3] - P[0]
        vol = abs(cc[0] * (a[1] * b[2] - a[2] * b[1]) + cc[1] * (a[2] * b[0] - a[0] * b[2]) + cc[2] * (a[0] * b[1] - a[1] * b[0])) / 6.0
        assert float(m.cell_volumes[c]) == vol
        R = basis_tet_reading(P, vol)
        assert np.array_equal(B[:, :, c], np.array(R)), c
        psi = P @ B[:3, :, c] + B[3, :, c]                 # psi[node j, function i]
        assert np.allclose(psi, np.eye(4), atol=1e-9)
    w.species = sp
    w.cfg.pic_type = capi.PIC_EXPLICIT
    w.cfg.collision_type = capi.NO_COLL
    w.cfg.fnum, w.cfg.eps_scaling = 2.5e6, 1.3
    ic = 18
    lam = np.array([0.1, 0.2, 0.3, 0.4])
    xp = [float(v) for v in xyz[:, cn[:, ic - 1]] @ lam]
    p = wl.new_particles(2)
    p["X"], p["Y"], p["Z"] = xp[0], xp[1], xp[2]
    p["S_ID"] = [sp.id("Ar"), sp.id("Ar+")]                 # the neutral is skipped
    p["IC"] = ic
    p["DTRIM"] = w.cfg.dt
    w.particles = p
    wl.apply(oracle, w)
    rhs = oracle.deposit_charge(m.nn)
    K = QE / (EPS0 * w.cfg.eps_scaling ** 2)
    rho_q = K * float(sp.charge[sp.id("Ar+") - 1]) * w.cfg.fnum
    want = np.zeros(m.nn)
    R = basis_tet_reading(xyz[:, cn[:, ic - 1]].T, float(m.cell_volumes[ic - 1]))
    for pp in range(4):
        psip = R[0][pp] * xp[0] + R[1][pp] * xp[1] + R[2][pp] * xp[2] + R[3][pp]
        want[cn[pp, ic - 1]] = 0.0 + rho_q * psip
    assert np.array_equal(rhs, want) and np.count_nonzero(rhs) == 4
    assert np.allclose(want[cn[:, ic - 1]] / rho_q, lam, atol=1e-9)          # the weights are the barycentric coordinates
    # ---- segments
    w = wl.c4_ccp_1d(ncells=9, n_per_species=1)
    m = w.mesh
    x = np.asarray(m.node_coords)[0]
    cn = np.asarray(m.cell_nodes) - 1
    B = np.asarray(m.basis_coeffs)
    for c in range(m.nc):
        x1, x2 = float(x[cn[0, c]]), float(x[cn[1, c]])
        length = abs(x2 - x1)
        want_b = [[-1.0 / length, 1.0 / length], [x2 / length, -x1 / length]]      # BASIS_COEFFS(j, i), :970-977
        assert np.array_equal(B[:, :, c], np.array(want_b)), c
    w.cfg.collision_type = capi.NO_COLL
    w.cfg.fnum = 4.0e9
    w.cfg.ymin, w.cfg.ymax, w.cfg.zmin, w.cfg.zmax = 0.0, 0.02, 0.0, 0.05
    ic = 4
    xp = float(0.3 * x[cn[0, ic - 1]] + 0.7 * x[cn[1, ic - 1]])
    p = w.particles[:1].copy()
    p["X"], p["IC"], p["S_ID"] = xp, ic, w.species.id("e")
    w.particles = p
    w.efield = None
    wl.apply(oracle, w)
    rhs = oracle.deposit_charge(m.nn)
    rho_q = (QE / (EPS0 * 1.0 ** 2)) * float(w.species.charge[w.species.id("e") - 1]) * w.cfg.fnum / (w.cfg.ymax - w.cfg.ymin) / (w.cfg.zmax - w.cfg.zmin)
    want = np.zeros(m.nn)
    for pp in range(2):
        want[cn[pp, ic - 1]] = 0.0 + rho_q * (float(B[0, pp, ic - 1]) * xp + float(B[1, pp, ic - 1]))
    assert np.array_equal(rhs, want) and np.count_nonzero(rhs) == 2


# ---- dielectric surface charging ---------------------------------------------------------------------------------------------
def test_dielectric_surface_charge_oracle_matches_fortran_reading(oracle):
    """An ion and an electron reach the face z = 0 of the pin tetrahedron, whose group is a DIELECTRIC_BC for the field and a
    specular wall for the particles: SURFACE_CHARGE(VP) += K*CHARGE*FNUM * PSI_P(hit point) on the four nodes of the cell
    (src/timecycle.f90:1398-1430), PSI from BASIS_COEFFS as the Fortran evaluates them; a neutral leaves no charge."""
    import test_fortran_pins as t1
    QE, EPS0 = 1.602176634e-19, 8.8541878128e-12
    mesh = t1.PINS["tet_mesh"]
    m = t1.pin_tet_mesh()
    nodes, cells = np.array(mesh["nodes"]), np.array(mesh["cells"])
    B = np.zeros((4, 4, m.nc))
    for c in range(m.nc):
        B[:, :, c] = np.array(basis_tet_reading(nodes[cells[c] - 1], float(m.cell_volumes[c])))
    m.basis_coeffs = np.asfortranarray(B)
    sp = wl.load_species("ar", ["Ar+", "e", "Ar"])
    dt = 1.0e-3
    cfg = capi.make_config(dims=3, grid_type=capi.UNSTRUCTURED, nx=1, ny=1, xmin=0.0, xmax=2.0, ymin=0.0, ymax=2.0, zmin=0.0, zmax=2.0,
                           dt=dt, fnum=6.0e7, collision_type=capi.NO_COLL, dump_bound_avg_every=0, eps_scaling=1.5)
    starts = [("Ar+", [0.3, 0.2, 0.2], [210.0, -95.0, -480.0]), ("e", [0.25, 0.3, 0.15], [-130.0, 60.0, -520.0]),
              ("Ar", [0.2, 0.25, 0.1], [50.0, 40.0, -400.0])]
    p = wl.new_particles(3)
    for k, (name, x, v) in enumerate(starts):
        p["X"][k], p["Y"][k], p["Z"][k] = x
        p["VX"][k], p["VY"][k], p["VZ"][k] = v
        p["S_ID"][k] = sp.id(name)
    p["IC"] = 1
    p["DTRIM"] = dt
    bcs = dict(particle_bc=[capi.SPECULAR] * 4, field_bc=[capi.DIELECTRIC_BC, capi.NO_BC, capi.NO_BC, capi.NO_BC],
               wall_temp=[300.0] * 4, acc_n=[1.0] * 4, acc_t=[1.0] * 4)
    w = wl.Workload("pin dielectric", cfg, sp, p, mesh=m, bcs=bcs)
    wl.apply(oracle, w)
    oracle.advect(1)
    got = np.zeros(m.nn)
    oracle._call("get_surface_charge", got.ctypes.data_as(__import__("ctypes").c_void_p), __import__("ctypes").c_int64(m.nn))
    want = np.zeros(m.nn)
    K = QE / (EPS0 * cfg.eps_scaling ** 2)
    for name, x, v in starts:
        ic = 1
        # first crossing: the face loop of src/timecycle.f90:1257-1305 in cell 1
        dtcoll, boundcoll = dt, -1
        for i in range(1, 5):
            n = mesh["normal"][ic - 1][i - 1]
            P = mesh["nodes"][mesh["cells"][ic - 1][i - 1] - 1]
            vn = v[0] * n[0] + v[1] * n[1] + v[2] * n[2]
            dx = (P[0] - x[0]) * n[0] + (P[1] - x[1]) * n[1] + (P[2] - x[2]) * n[2]
            if vn >= 0.0 and vn * dtcoll >= dx:
                dtcoll, boundcoll = dx / vn, i
        assert boundcoll == 1 and mesh["faces_pg"][0][0] == 1
        hit = [x[k] + v[k] * dtcoll for k in range(3)]
        charge = float(sp.charge[sp.id(name) - 1])
        if abs(charge) >= 1.0e-6:
            rho_q = K * charge * cfg.fnum
            for i in range(4):
                vp = cells[ic - 1][i]
                psip = B[0, i, ic - 1] * hit[0] + B[1, i, ic - 1] * hit[1] + B[2, i, ic - 1] * hit[2] + B[3, i, ic - 1]
                want[vp - 1] = want[vp - 1] + rho_q * psip
    assert np.count_nonzero(want) >= 3 and np.array_equal(got, want), (got, want)   # (the basis function of the opposite node vanishes on the face)


# ---- rectilinear axisymmetric grid ------------------------------------------------------------------------------------------
def test_structured_axisymmetric_walk_oracle_matches_fortran_reading(oracle):
    """Box sides of an axisymmetric rectilinear grid (the radius quadratic, roots >= 0 without the TOL of the unstructured
    test, trial moves), a cone-shaped and a disc-shaped interior wall, all specular: bit-identical with the plain-Python
    reading of src/timecycle.f90:1662-1729, 1762-1838, 1885-1957."""
    t = PINS["structured_axi"]
    g = t["grid"]
    sp = wl.load_species("air", ["N2", "O2"])
    dt = fh(t["dt"])
    cfg = capi.make_config(dims=2, axi=1, grid_type=capi.RECTILINEAR_UNIFORM, nx=g["nx"], ny=g["ny"], xmin=g["xmin"], xmax=g["xmax"],
                           ymin=g["ymin"], ymax=g["ymax"], zmin=g["zmin"], zmax=g["zmax"], dt=dt, fnum=1.0, collision_type=capi.NO_COLL,
                           bool_periodic=g["periodic"], bool_specular=g["specular"], bool_diffuse=g["diffuse"], boundtemp=g["boundtemp"],
                           bool_z_periodic=g["z_periodic"])
    cases = t["cases"]
    p = wl.new_particles(len(cases))
    ddx, ddy = (g["xmax"] - g["xmin"]) / g["nx"], (g["ymax"] - g["ymin"]) / g["ny"]
    for k, c in enumerate(cases):
        q = c["inp"]
        for f in ("x", "y", "z", "vx", "vy", "vz"):
            p[f.upper()][k] = fh(q[f])
        p["IC"][k] = int((fh(q["x"]) - g["xmin"]) / ddx) + g["nx"] * int((fh(q["y"]) - g["ymin"]) / ddy) + 1
    p["S_ID"] = 1
    p["DTRIM"] = dt
    ws = t["walls"]
    col = lambda k: [w[k] for w in ws]
    walls = dict(x1=col("x1"), y1=col("y1"), x2=col("x2"), y2=col("y2"), normx=col("normx"), normy=col("normy"), temp=col("temp"),
                 specular=col("specular"), diffuse=col("diffuse"), porous=col("porous"), transmissivity=col("transm"), react=[0] * len(ws))
    w = wl.Workload("pin axisymmetric rectilinear grid", cfg, sp, p, walls=walls)
    w.cfg.track_id = 1
    wl.apply(oracle, w)
    oracle.advect(1)
    out = util.by_id(oracle.download_particles())
    assert len(out) == len(cases)
    for k, c in enumerate(cases):
        e = c["out"]
        assert out["IC"][k] == e["ic"], (k, out["IC"][k], e["ic"], c["events"])
        for f in ("x", "y", "z", "vx", "vy", "vz"):
            got = float(out[f.upper()][k])
            assert got == fh(e[f]), (k, f, got.hex(), e[f], c["events"])


# ---- plane triangles with a SOLID block --------------------------------------------------------------------------------------
def test_plane_triangle_walk_with_a_solid_block_oracle_matches_fortran_reading(oracle):
    """The DIMS == 2 face test and the FLUIDBOUNDARY rule: a face towards a cell of a SOLID volume group is a wall with the
    model of the face's own physical group (src/timecycle.f90:1270-1285, 1327-1338), hops, specular walls: bit-identical."""
    t, ts = PINS["tri_walk_solid"], PINS["tri_solid"]
    m = pin_tri_mesh()
    m.cell_faces_pg = np.asfortranarray(np.array(ts["faces_pg"], dtype=np.int32).T)
    m.cell_pg = np.array(ts["cell_pg"], dtype=np.int32)
    xy = np.array(PINS["tri_mesh"]["nodes"])
    tri = np.array(PINS["tri_mesh"]["cells"]) - 1
    a, b, c = xy[tri[:, 0]], xy[tri[:, 1]], xy[tri[:, 2]]
    m.cell_volumes = 0.5 * np.abs(a[:, 0] * (b[:, 1] - c[:, 1]) + b[:, 0] * (c[:, 1] - a[:, 1]) + c[:, 0] * (a[:, 1] - b[:, 1]))
    sp = wl.load_species("air", ["N2", "O2"])
    dt = fh(t["dt"])
    cfg = capi.make_config(dims=2, grid_type=capi.UNSTRUCTURED, nx=1, ny=1, xmin=0.0, xmax=0.09, ymin=0.01, ymax=0.05, zmin=-0.5, zmax=0.5,
                           dt=dt, fnum=1.0, collision_type=capi.NO_COLL, dump_bound_avg_every=0, bool_z_periodic=0)
    p = wl.new_particles(len(t["cases"]))
    for k, cse in enumerate(t["cases"]):
        q = cse["inp"]
        for f in ("x", "y", "z", "vx", "vy", "vz"):
            p[f.upper()][k] = fh(q[f])
        p["IC"][k] = q["ic"]
    p["S_ID"] = 1
    p["DTRIM"] = dt
    nb = 6
    volume_bc = [capi.FLUID] * nb
    volume_bc[ts["solid_group"] - 1] = capi.SOLID
    bcs = dict(particle_bc=[capi.SPECULAR] * nb, volume_bc=volume_bc, wall_temp=[300.0] * nb, acc_n=[1.0] * nb, acc_t=[1.0] * nb)
    w = wl.Workload("pin plane triangles + solid block", cfg, sp, p, mesh=m, bcs=bcs)
    w.cfg.track_id = 1
    wl.apply(oracle, w)
    oracle.advect(1)
    out = util.by_id(oracle.download_particles())
    assert len(out) == len(t["cases"])
    assert sum(1 for cse in t["cases"] for (cc, f, fb) in cse["trace"] if fb and PINS["tri_mesh"]["neighbors"][cc - 1][f - 1] != -1) >= 2
    for k, cse in enumerate(t["cases"]):
        e = cse["out"]
        assert out["IC"][k] == e["ic"], (k, out["IC"][k], e["ic"], cse["trace"])
        for f in ("x", "y", "z", "vx", "vy", "vz"):
            got = float(out[f.upper()][k])
            assert got == fh(e[f]), (k, f, got.hex(), e[f], cse["trace"])


# ---- BOUNDARIES_EMIT on 2-D meshes -------------------------------------------------------------------------------------------
@pytest.mark.parametrize("label", ["plane", "axi"])
def test_boundaries_emit_2d_oracle_matches_fortran_reading(oracle, label):
    """src/timecycle.f90:368-547 (DIMS == 2) on scripted draws: the tangent (-n_y, n_x, 0), the point on the edge (plane: linear,
    Z drawn; axisymmetric: area-weighted radius with Y1 < Y2, Y1 > Y2 and Y1 == Y2), creation order over tasks / components."""
    t = PINS["emit_2d"][label]
    axi = label == "axi"
    m = pin_tri_mesh()
    if not axi:
        xy = np.array(PINS["tri_mesh"]["nodes"])
        tri = np.array(PINS["tri_mesh"]["cells"]) - 1
        a, b, c = xy[tri[:, 0]], xy[tri[:, 1]], xy[tri[:, 2]]
        m.cell_volumes = 0.5 * np.abs(a[:, 0] * (b[:, 1] - c[:, 1]) + b[:, 0] * (c[:, 1] - a[:, 1]) + c[:, 0] * (a[:, 1] - b[:, 1]))
    sp = wl.load_species("air", ["N2", "O2"])
    cfg = capi.make_config(dims=2, axi=int(axi), grid_type=capi.UNSTRUCTURED, nx=1, ny=1, xmin=0.0, xmax=0.09, ymin=0.01, ymax=0.05,
                           zmin=t["zmin"], zmax=t["zmax"], dt=fh(t["dt"]), fnum=1.0, collision_type=capi.NO_COLL, dump_bound_avg_every=0)
    cfg.capacity = 256
    bcs = dict(particle_bc=[capi.SPECULAR] * 4, wall_temp=[300.0] * 4, acc_n=[1.0] * 4, acc_t=[1.0] * 4)
    w = wl.Workload("pin emit 2-D", cfg, sp, wl.new_particles(0), mesh=m, bcs=bcs)
    wl.apply(oracle, w)
    oracle.set_mixtures([[(sp.id(name), frac) for name, frac in t["mix"]]])
    oracle.set_species_weights([1.0] * len(sp.mass))
    oracle.set_emit_tasks([dict(task, mix_id=0) for task in t["tasks"]])
    script(oracle, [fh(v) for v in t["draws"]])
    oracle.boundaries_emit(1)
    assert script_used(oracle) == t["used"]
    out = oracle.download_particles()
    assert len(out) == len(t["out"]) >= 9
    for k, e in enumerate(t["out"]):
        assert out["S_ID"][k] == sp.id(e["sid"]) and out["IC"][k] == e["ic"], k
        for f in ("x", "y", "z", "vx", "vy", "vz", "erot", "evib", "dtrim"):
            assert float(out[f.upper()][k]) == pytest.approx(fh(e[f]), rel=1e-13, abs=1e-300), (k, f)


# ---- INITIAL_SEED on triangles and segments ----------------------------------------------------------------------------------
@pytest.mark.parametrize("label", ["plane", "axi", "line"])
def test_initial_seed_2d_1d_oracle_matches_fortran_reading(oracle, label):
    """src/initialization.f90:2062-2141 on scripted draws: RANDINT per (task, component, cell), the (S, T) fold, the point in
    the triangle (Z drawn, or 0 when axisymmetric -- U is drawn either way) or on the segment (Y from T, Z from U), MAXWELL
    with a zero temperature component, INTERNAL_ENERGY with a zero temperature, SPWT; creation order."""
    t = PINS["seed_2d_1d"][label]
    d = t["dims"]
    ymin, ymax, zmin, zmax = t["lim"]
    nn, nc = len(t["nodes"]), len(t["cells"])
    if d == 2:
        m = pin_tri_mesh()
    else:
        from pantera_pic_dsmc_b200 import meshgen as mg
        m = mg.line_1d(nc, 0.0, 1.0)
        xyz = np.zeros((3, nn))
        xyz[0] = [n[0] for n in t["nodes"]]
        m.node_coords = np.asfortranarray(xyz)
        assert [list(c) for c in (np.asarray(m.cell_nodes).T)] == t["cells"]
    m.cell_volumes = np.array([fh(v) for v in t["vols"]])
    sp = wl.load_species("air", ["N2", "O2"])
    xs = [n[0] for n in t["nodes"]]
    cfg = capi.make_config(dims=d, axi=int(t["axi"]), grid_type=capi.UNSTRUCTURED, nx=1, ny=1, xmin=min(xs), xmax=max(xs), ymin=ymin, ymax=ymax,
                           zmin=zmin, zmax=zmax, dt=1.0e-6, fnum=fh(t["fnum"]), collision_type=capi.NO_COLL, dump_bound_avg_every=0)
    cfg.capacity = 512
    nb = 4 if d == 2 else 2
    bcs = dict(particle_bc=[capi.SPECULAR] * nb, wall_temp=[300.0] * nb, acc_n=[1.0] * nb, acc_t=[1.0] * nb)
    w = wl.Workload("pin seed", cfg, sp, wl.new_particles(0), mesh=m, bcs=bcs)
    wl.apply(oracle, w)
    oracle.set_mixtures([[(sp.id(name), frac) for name, frac in mix] for mix in t["mixtures"]])
    oracle.set_species_weights([t["spwt"][n] for n in ("N2", "O2")])
    oracle.set_initial_particles_tasks([dict(q, rate=q["nrho"]) for q in t["tasks"]])
    script(oracle, [fh(v) for v in t["draws"]])
    oracle.initial_seed()
    assert script_used(oracle) == t["used"]
    out = oracle.download_particles()
    assert len(out) == len(t["out"]) >= 10
    for k, e in enumerate(t["out"]):
        assert out["S_ID"][k] == sp.id(e["sid"]) and out["IC"][k] == e["ic"], k
        for f in ("x", "y", "z", "vx", "vy", "vz", "erot", "evib"):
            assert float(out[f.upper()][k]) == pytest.approx(fh(e[f]), rel=1e-13, abs=1e-300), (k, f)


def test_volume_inject_on_triangles_oracle_matches_fortran_reading(oracle):
    """src/timecycle.f90:553-670 (unstructured branch) on scripted draws: RANDINT(DT*NRHODOT/(FNUM*SPWT)*VOL*MOLFRAC/N) per
    (task, component, cell), the triangle point, MAXWELL with a drift, INTERNAL_ENERGY (TVIB = 0: no draw), DTRIM = DT."""
    t = PINS["vinject_tri"]
    ymin, ymax, zmin, zmax = t["lim"]
    m = pin_tri_mesh()
    m.cell_volumes = np.array([fh(v) for v in t["vols"]])
    sp = wl.load_species("air", ["N2", "O2"])
    dt = fh(t["dt"])
    cfg = capi.make_config(dims=2, grid_type=capi.UNSTRUCTURED, nx=1, ny=1, xmin=0.0, xmax=0.09, ymin=ymin, ymax=ymax, zmin=zmin, zmax=zmax,
                           dt=dt, fnum=1.0, collision_type=capi.NO_COLL, dump_bound_avg_every=0)
    cfg.capacity = 512
    bcs = dict(particle_bc=[capi.SPECULAR] * 4, wall_temp=[300.0] * 4, acc_n=[1.0] * 4, acc_t=[1.0] * 4)
    w = wl.Workload("pin volume inject", cfg, sp, wl.new_particles(0), mesh=m, bcs=bcs)
    wl.apply(oracle, w)
    oracle.set_mixtures([[(sp.id(name), frac) for name, frac in mix] for mix in t["mixtures"]])
    oracle.set_species_weights([t["spwt"][n] for n in ("N2", "O2")])
    oracle.set_volume_inject_tasks([dict(q, rate=q["nrhodot"]) for q in t["tasks"]])
    script(oracle, [fh(v) for v in t["draws"]])
    oracle.volume_inject(3)
    assert script_used(oracle) == t["used"]
    out = oracle.download_particles()
    assert len(out) == len(t["out"]) >= 10
    for k, e in enumerate(t["out"]):
        assert out["S_ID"][k] == sp.id(e["sid"]) and out["IC"][k] == e["ic"], k
        for f in ("x", "y", "z", "vx", "vy", "vz", "erot", "evib"):
            assert float(out[f.upper()][k]) == pytest.approx(fh(e[f]), rel=1e-13, abs=1e-300), (k, f)
        assert float(out["DTRIM"][k]) == dt


# ---- REMOVE_PARTICLES_OF_SPECIES ----------------------------------------------------------------------------------------------
def test_remove_particles_of_species_store_order_matches_fortran_reading(oracle):
    """src/timecycle.f90:2340-2367 with REMOVE_PARTICLE_ARRAY (src/particle.f90:142-179): reverse sweep, the last particle takes
    the place of a removed one -- the surviving ORDER is what the next sequential MT draws act on."""
    w = wl.c2_air_box(40, nx=2, ny=2, with_reactions=False)
    rng = np.random.default_rng(3)
    p = w.particles
    p["S_ID"] = rng.integers(1, 4, len(p))
    p["ID"] = 500 + np.arange(len(p))
    w.particles = p
    wl.apply(oracle, w)
    store = [(int(i), int(s)) for i, s in zip(p["ID"], p["S_ID"])]
    for sp_id in (2, 3):                                   # REMOVE_PARTICLES_IN_MIXTURE: component after component
        oracle.remove_species(sp_id)
        ip = len(store)
        while ip >= 1:
            if store[ip - 1][1] == sp_id:
                store[ip - 1] = store[-1]
                store.pop()
            ip -= 1
        got = oracle.download_particles()
        assert [int(i) for i in got["ID"]] == [i for i, _ in store]
    assert 0 < len(store) < 40 and all(s == 1 for _, s in store)


# ---- diffuse wall on plane triangles --------------------------------------------------------------------------------------------
def test_diffuse_wall_on_triangles_oracle_matches_fortran_reading(oracle):
    """src/timecycle.f90:1327-1343, 1445-1476 on scripted draws: the 2-D wall frame (FACE_NORMAL = -EDGE_NORMAL, FACE_TANG1 =
    (-n_y, n_x, 0), FACE_TANG2 = z) and the re-emission V = V_PERP n + V_TANG1 t1 + V_TANG2 t2 with MAXWELL / INTERNAL_ENERGY /
    FLX at the wall temperature of the face's group; two particles share the stream in store order."""
    t = PINS["tri_diffuse"]
    m = pin_tri_mesh()
    xy = np.array(PINS["tri_mesh"]["nodes"])
    tri = np.array(PINS["tri_mesh"]["cells"]) - 1
    a, b, c = xy[tri[:, 0]], xy[tri[:, 1]], xy[tri[:, 2]]
    m.cell_volumes = 0.5 * np.abs(a[:, 0] * (b[:, 1] - c[:, 1]) + b[:, 0] * (c[:, 1] - a[:, 1]) + c[:, 0] * (a[:, 1] - b[:, 1]))
    sp = wl.load_species("air", ["N2", "O2"])
    dt = fh(t["dt"])
    cfg = capi.make_config(dims=2, grid_type=capi.UNSTRUCTURED, nx=1, ny=1, xmin=0.0, xmax=0.09, ymin=0.01, ymax=0.05, zmin=-0.5, zmax=0.5,
                           dt=dt, fnum=1.0, collision_type=capi.NO_COLL, dump_bound_avg_every=0, bool_z_periodic=0)
    p = wl.new_particles(len(t["cases"]))
    for k, cse in enumerate(t["cases"]):
        q = cse["inp"]
        for f in ("x", "y", "z", "vx", "vy", "vz", "erot", "evib"):
            p[f.upper()][k] = fh(q[f])
        p["IC"][k] = q["ic"]
        p["S_ID"][k] = sp.id(q["sid"])
    p["DTRIM"] = dt
    kinds = [capi.SPECULAR] * 4
    temps = [300.0] * 4
    for g, tw in t["walls"].items():
        kinds[int(g) - 1] = capi.DIFFUSE
        temps[int(g) - 1] = tw
    bcs = dict(particle_bc=kinds, wall_temp=temps, acc_n=[1.0] * 4, acc_t=[1.0] * 4)
    w = wl.Workload("pin diffuse triangles", cfg, sp, p, mesh=m, bcs=bcs)
    w.cfg.track_id = 1
    wl.apply(oracle, w)
    script(oracle, [fh(v) for v in t["draws"]])
    oracle.advect(1)
    assert script_used(oracle) == t["used"]
    out = util.by_id(oracle.download_particles())
    for k, cse in enumerate(t["cases"]):
        e = cse["out"]
        assert out["IC"][k] == e["ic"], (k, cse["trace"])
        for f in ("x", "y", "z", "vx", "vy", "vz", "erot", "evib"):
            assert float(out[f.upper()][k]) == pytest.approx(fh(e[f]), rel=1e-13, abs=1e-300), (k, f, cse["trace"])


# ---- 1-D line mesh: the C4 geometry --------------------------------------------------------------------------------------------
def test_line_mesh_walk_oracle_matches_fortran_reading(oracle):
    """ADVECT on a 1-D unstructured mesh in plain Python: the DIMS == 1 face test (src/timecycle.f90:1257-1269), hops, a specular
    electrode on the right, a vacuum one on the left (the particle is removed: :1611-1614), then the Y and Z periodic wraps
    of 1-D runs (:2044-2066) -- bit-identical, survivors matched by ID (the removal sweep reorders the store)."""
    xs = [0.0, 0.011, 0.019, 0.034, 0.041, 0.05]
    nc = len(xs) - 1
    from pantera_pic_dsmc_b200 import meshgen as mg
    m = mg.line_1d(nc, 0.0, 1.0)
    xyz = np.zeros((3, nc + 1))
    xyz[0] = xs
    m.node_coords = np.asfortranarray(xyz)
    # EDGE_NORMAL of a segment: -1 at its first node, +1 at its second (src/grid_and_partition.f90:931-937); neighbours along the line
    normal = [(-1.0, 1.0)] * nc
    assert np.array_equal(np.asarray(m.face_normal)[0], np.array(normal).T)
    sp = wl.load_species("ar", ["Ar", "Ar+", "e"])
    dt = 4.0e-9
    ymin, ymax, zmin, zmax = 0.0, 0.02, -0.01, 0.03
    cfg = capi.make_config(dims=1, grid_type=capi.UNSTRUCTURED, nx=1, ny=1, xmin=0.0, xmax=0.05, ymin=ymin, ymax=ymax, zmin=zmin, zmax=zmax,
                           dt=dt, fnum=1.0, collision_type=capi.NO_COLL, dump_bound_avg_every=0, bool_z_periodic=1, bool_y_periodic=1)
    starts = [(0.013, 0.004, 0.0, 6.0e6, 9.0e6, -2.5e7),      # to the right: hops, specular electrode, back; y and z wrap several times
              (0.03, 0.019, 0.02, -9.0e6, 1.0e6, 4.0e6),       # to the left: reaches x = 0 -> removed
              (0.045, 0.01, 0.0, 1.0e5, -3.0e6, 1.0e6),         # stays in its cell; y wraps downwards
              (0.02, 0.0, 0.029, -2.0e6, 0.0, 2.0e6)]           # one hop to the left, z wraps once
    p = wl.new_particles(len(starts))
    for k, (x, y, z, vx, vy, vz) in enumerate(starts):
        p["X"][k], p["Y"][k], p["Z"][k], p["VX"][k], p["VY"][k], p["VZ"][k] = x, y, z, vx, vy, vz
        p["IC"][k] = max(i for i in range(nc) if xs[i] <= x) + 1
    p["S_ID"] = sp.id("Ar")
    p["DTRIM"] = dt
    p["ID"] = 70 + np.arange(len(starts))
    bcs = dict(particle_bc=[capi.VACUUM, capi.SPECULAR], wall_temp=[300.0] * 2, acc_n=[1.0] * 2, acc_t=[1.0] * 2)
    w = wl.Workload("pin line mesh", cfg, sp, p, mesh=m, bcs=bcs)
    wl.apply(oracle, w)
    oracle.advect(1)
    got = {int(q["ID"]): q for q in oracle.download_particles()}
    expect = {}
    for k, (x, y, z, vx, vy, vz) in enumerate(starts):
        ic = int(p["IC"][k])
        dtrim = dt
        removed = False
        while dtrim > 0.0:
            dtcoll, boundcoll = dtrim, -1
            for i in (1, 2):
                n = normal[ic - 1][i - 1]
                vn = vx * n
                dx = (xs[ic - 1 + (i - 1)] - x) * n
                if vn >= 0.0 and vn * dtcoll >= dx:
                    dtcoll, boundcoll = dx / vn, i
            if boundcoll == -1:
                x, y, z = x + vx * dtrim, y + vy * dtrim, z + vz * dtrim
                dtrim = 0.0
                continue
            x, y, z = x + vx * dtcoll, y + vy * dtcoll, z + vz * dtcoll
            dtrim = dtrim - dtcoll
            neighbor = ic - 1 if boundcoll == 1 else ic + 1
            if 1 <= neighbor <= nc:
                ic = neighbor
            elif boundcoll == 2:                                 # right electrode, group 2: specular
                fn = -normal[ic - 1][1]
                vdotn = vx * fn + vy * 0.0 + vz * 0.0
                vx, vy, vz = vx - 2.0 * vdotn * fn, vy - 2.0 * vdotn * 0.0, vz - 2.0 * vdotn * 0.0
            else:                                                # left electrode, group 1: vacuum
                removed = True
                dtrim = 0.0
        while z > zmax:
            z = zmin + (z - zmax)
        while z < zmin:
            z = zmax + (z - zmin)
        while y > ymax:
            y = ymin + (y - ymax)
        while y < ymin:
            y = ymax + (y - ymin)
        if not removed:
            expect[70 + k] = (x, y, z, vx, vy, vz, ic)
    assert sorted(got) == sorted(expect) == [70, 72, 73]
    for i, e in expect.items():
        q = got[i]
        assert (float(q["X"]), float(q["Y"]), float(q["Z"]), float(q["VX"]), float(q["VY"]), float(q["VZ"]), int(q["IC"])) == e, (i, e)
