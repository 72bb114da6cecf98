"""Restart re-binning: REASSIGN_PARTICLES_TO_CELLS_1D / _2D / _3D (src/tools.f90:925-1039, 782-922,
577-778), called by PROGRAM PANTERA right after READ_PARTICLES_FILE (src/pantera.f90:58-70).

* `fortran_reassign` below is an independent plain-Python reading of the Fortran (explicit loops, the reference's own
  bin lists in ascending cell order, first INSIDE cell wins) -- not derived from oracle/pantera_oracle.c; the oracle must
  give the same IC for every particle and remove the same ones;
* the device must agree with the oracle bit for bit (IC is an integer result) on line, triangle and tetrahedral meshes,
  including particles exactly on shared faces and nodes (the lowest cell id wins) and particles outside the domain box
  or outside every cell (removed);
* the restart flow: a dump written on one mesh is read into a context on ANOTHER mesh with its IC untrusted
  (PGPU_RESTART_UNASSIGNED); the hot path refuses to run until the particles have been located again.
"""
import numpy as np
import pytest

from pantera_pic_dsmc_b200 import capi, meshgen as mg, workloads as wl

import util

NPART = 100  # N_PARTITIONS_X/Y/Z


def fortran_index(c, lo, h):
    """INT((C - XMIN)/DX) + 1, clamped to 1..100 where the Fortran would leave its arrays."""
    q = (c - lo) / h
    if not q < NPART:
        return NPART
    return max(int(q) + 1, 1)  # int() truncates toward zero like INT()


def fortran_reassign(mesh, cfg, parts):
    """Returns (ic per particle or 0 when the particle is removed)."""
    d = mesh.dims
    lo = [cfg.xmin, cfg.ymin, cfg.zmin][:d]
    hi = [cfg.xmax, cfg.ymax, cfg.zmax][:d]
    h = [(hi[k] - lo[k]) / NPART for k in range(d)]
    xyz = np.asarray(mesh.node_coords)          # (3, nn)
    cn = np.asarray(mesh.cell_nodes)            # (d+1, nc), 1-based
    bc = np.asarray(mesh.basis_coeffs)          # (d+1, d+1, nc): BASIS_COEFFS(j, vp, ic)
    bins = {}
    for ic in range(1, mesh.nc + 1):            # DO IC = 1, NCELLS: lists fill in ascending cell order
        rng = []
        for k in range(d):
            xs = [xyz[k, cn[p, ic - 1] - 1] for p in range(d + 1)]
            rng.append(range(fortran_index(min(xs), lo[k], h[k]), fortran_index(max(xs), lo[k], h[k]) + 1))
        while len(rng) < 3:
            rng.append(range(1, 2))
        for i in rng[0]:
            for j in rng[1]:
                for k in rng[2]:
                    bins.setdefault((i, j, k), []).append(ic)
    out = np.zeros(len(parts), dtype=np.int32)
    for n, q in enumerate(parts):
        xp = [q["X"], q["Y"], q["Z"]][:d]
        if any(xp[k] < lo[k] or xp[k] > hi[k] for k in range(d)):
            continue
        idx = [fortran_index(xp[k], lo[k], h[k]) for k in range(d)] + [1] * (3 - d)
        for ic in bins.get(tuple(idx), []):
            inside = True
            for vp in range(d + 1):
                psip = xp[0] * bc[0, vp, ic - 1]
                for k in range(1, d):
                    psip = psip + xp[k] * bc[k, vp, ic - 1]
                psip = psip + bc[d, vp, ic - 1]
                if psip < 0:
                    inside = False
            if inside:
                out[n] = ic
                break
    return out


def scrambled(w, seed, extra_outside=40):
    """The workload's particles with every IC scrambled, some particles moved onto mesh nodes and edge midpoints (ties
    between cells), some pushed outside the domain box."""
    rng = np.random.default_rng(seed)
    p = w.particles.copy()
    m = w.mesh
    d = m.dims
    xyz = np.asarray(m.node_coords)
    cn = np.asarray(m.cell_nodes) - 1
    n = len(p)
    k = min(n // 10, 300)
    nodes = rng.integers(0, m.nn, k)                         # exactly on a node
    for f, row in zip("XYZ"[:d], range(d)):
        p[f][:k] = xyz[row, nodes]
    cells = rng.integers(0, m.nc, k)                         # exactly halfway between the first two nodes of a cell
    for f, row in zip("XYZ"[:d], range(d)):
        p[f][k:2 * k] = 0.5 * (xyz[row, cn[0, cells]] + xyz[row, cn[1, cells]])
    out = slice(2 * k, 2 * k + extra_outside)
    p["X"][out] = np.where(rng.random(extra_outside) < 0.5, w.cfg.xmin - 1e-3, w.cfg.xmax + 1e-3)
    p["IC"] = rng.integers(1, m.nc + 1, n)
    return p


def workload(kind):
    if kind == "line":
        w = wl.c4_ccp_1d(37, 1500)
        w.particles = w.particles[w.particles["S_ID"] > 0]
    elif kind == "tris":
        w = wl.c3_tri_box(nx=9, ny=7, n=6000)
    elif kind == "tris-jitter":
        w = wl.c3_tri_box(nx=9, ny=7, n=6000)
        rng = np.random.default_rng(2)
        nx, ny = 9, 7
        X, Y = np.meshgrid(np.linspace(w.cfg.xmin, w.cfg.xmax, nx + 1), np.linspace(w.cfg.ymin, w.cfg.ymax, ny + 1), indexing="ij")
        X[1:-1, 1:-1] += (rng.random((nx - 1, ny - 1)) - 0.5) * 0.4 * (w.cfg.xmax - w.cfg.xmin) / nx
        Y[1:-1, 1:-1] += (rng.random((nx - 1, ny - 1)) - 0.5) * 0.4 * (w.cfg.ymax - w.cfg.ymin) / ny
        w.mesh = mg.mapped_quads_to_tris(X, Y, zmin=w.cfg.zmin, zmax=w.cfg.zmax)
    else:
        w = wl.c5_channel(5, 4, 3, ppc=12, jitter=0.15 if kind == "tets-jitter" else 0.0)
    w.cfg.track_id = 1
    return w


KINDS = ["line", "tris", "tris-jitter", "tets", "tets-jitter"]


@pytest.mark.parametrize("kind", KINDS)
def test_oracle_reassign_equals_the_plain_python_reading_of_the_fortran(oracle, kind):
    w = workload(kind)
    p = scrambled(w, 11)
    expect = fortran_reassign(w.mesh, w.cfg, p)
    assert (expect == 0).sum() >= 40 and (expect > 0).sum() > 0.9 * len(p) - 100
    w.particles = p
    wl.apply(oracle, w)
    removed = oracle.reassign_particles_to_cells()
    assert removed == int((expect == 0).sum())
    got = oracle.download_particles()
    # the swap-with-last removals reorder the store: match the survivors by their (untouched) velocities
    keep = expect > 0
    key = lambda a: np.lexsort((a["VZ"], a["VY"], a["VX"]))
    a, b = p[keep][key(p[keep])], got[key(got)]
    assert np.array_equal(a["VX"], b["VX"])
    assert np.array_equal(expect[keep][key(p[keep])], b["IC"])


@pytest.mark.parametrize("kind", KINDS)
def test_the_located_cell_contains_the_particle(oracle, kind):
    """Geometry check independent of any bin structure: all barycentric coordinates of the assigned cell are >= 0 and no
    lower-numbered cell of the whole mesh contains the particle."""
    w = workload(kind)
    p = scrambled(w, 5)
    w.particles = p
    wl.apply(oracle, w)
    oracle.reassign_particles_to_cells()
    got = oracle.download_particles()
    m = w.mesh
    d = m.dims
    bc = np.asarray(m.basis_coeffs)
    xp = np.stack([got["X"], got["Y"], got["Z"]])[:d]                       # (d, n)
    psi = np.einsum("kvc,kn->vcn", bc[:d], xp) + bc[d][:, :, None]           # (vp, cell, particle)
    inside = (psi >= -1e-12).all(axis=0)                                     # (cell, particle)
    assert inside[got["IC"] - 1, np.arange(len(got))].all()
    strictly = (psi > 1e-9).all(axis=0)
    first = np.argmax(strictly, axis=0) + 1
    has = strictly.any(axis=0)
    assert np.array_equal(first[has], got["IC"][has])                        # interior points: the unique containing cell


def device_vs_oracle(gpu, oracle, kind):
    w = workload(kind)
    w.particles = scrambled(w, 23)
    util.setup_pair(gpu, oracle, w)
    ro, rg = oracle.reassign_particles_to_cells(), gpu.reassign_particles_to_cells()
    assert ro == rg and ro >= 40
    a, b = util.by_id(oracle.download_particles()), util.by_id(gpu.download_particles())
    assert np.array_equal(a["ID"], b["ID"])
    util.assert_particles_match(a, b, exact_fields=util.FLOAT_FIELDS, what=f"re-binning on {kind}")


@pytest.mark.gpu
@pytest.mark.parametrize("kind", KINDS)
def test_device_reassign_is_bit_exact_with_the_oracle(gpu, oracle, kind):
    device_vs_oracle(gpu, oracle, kind)


@pytest.mark.gpu
def test_device_restart_onto_another_mesh(gpu, oracle, tmp_path):
    """Dump on a 9x7 triangulation, restart on a jittered 12x5 one: the IC of the file means nothing there."""
    w = workload("tris")
    wl.apply(gpu, w)
    prefix = str(tmp_path) + "/r_"
    gpu.dump_particles_file(prefix, 40)
    w2 = wl.c3_tri_box(nx=12, ny=5, n=10)
    w2.cfg.track_id = 1
    for lib in (gpu, oracle):
        wl.apply(lib, w2, upload=False)
        lib.upload_particles(w2.particles[:0])
    with pytest.raises(capi.PanteraError):               # strict mode: cell ids beyond the new mesh are an error
        gpu.read_particles_file(prefix, 40)
    gpu.upload_particles(w2.particles[:0])
    gpu.read_particles_file(prefix, 40, unassigned=True)
    oracle.read_particles_file(prefix, 40, unassigned=True)
    with pytest.raises(capi.PanteraError, match="reassign"):
        gpu.advect(1)                                    # the hot path refuses to run on unassigned cells
    assert gpu.reassign_particles_to_cells() == oracle.reassign_particles_to_cells() == 0
    a, b = oracle.download_particles(), gpu.download_particles()
    key = lambda q: np.lexsort((q["VZ"], q["VY"], q["VX"]))
    a, b = a[key(a)], b[key(b)]
    assert len(a) == len(w.particles) and np.array_equal(a["IC"], b["IC"]) and np.array_equal(a["X"], b["X"])
    assert a["IC"].max() <= w2.mesh.nc
    gpu.advect(1)                                        # and now it runs
    assert gpu.num_particles() > 0
