"""Synthetic simplicial meshes + the derived arrays the hot path consumes.

The reference ships no mesh; its readers ``READ_1D/2D/3D_UNSTRUCTURED_GRID_SU2``
(src/grid_and_partition.f90:627-2147) turn an SU2 file into the arrays of ``U1D/U2D/U3D_GRID``
(src/global.f90:128-193).  In a PANTERA run those arrays come from the Fortran host through
``pgpu_set_mesh``; here they are computed with the same formulas so that the
test/bench harness can feed both the oracle and the CUDA library:

* local numbering: 3-D face j = local nodes [1,3,2],[1,2,4],[2,3,4],[1,4,3] and local node j lies
  on face j (grid_and_partition.f90:1932-1936); 2-D edge j = nodes (j, j mod 3 + 1) with outward
  normal (y2-y1, x1-x2)/len (:1402-1418); 1-D face 1 = node 1 (normal -x), face 2 = node 2 (:931-937);
* volumes |C.(AxB)|/6 (:1845-1849), 2-D area*(ZMAX-ZMIN) [*rbar if axisymmetric] (:1315-1323),
  1-D len*(YMAX-YMIN)*(ZMAX-ZMIN) (:870-871);
* FACE_TANG1 = (B-A)/|B-A|, FACE_TANG2 = n x t1 (:1961-1962);
* P1 basis coefficients (:970-977, :1462-1474, :2017-2037).

All arrays have their *Fortran* shape (first index fastest) and ids are 1-based, -1 = none.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, Optional

import numpy as np

FACE_LOCAL_3D = np.array([[0, 2, 1], [0, 1, 3], [1, 2, 3], [0, 3, 2]])  # 0-based local nodes of face j
EDGE_LOCAL_2D = np.array([[0, 1], [1, 2], [2, 0]])


@dataclass
class Mesh:
    dims: int
    nn: int
    nc: int
    node_coords: np.ndarray  # (3, nn)
    cell_nodes: np.ndarray  # (d+1, nc) int32, 1-based
    cell_neighbors: np.ndarray  # (d+1, nc) int32
    face_normal: np.ndarray  # (3, d+1, nc)
    cell_faces_pg: np.ndarray  # (d+1, nc) int32
    cell_pg: np.ndarray  # (nc,) int32
    cell_volumes: np.ndarray  # (nc,)
    basis_coeffs: np.ndarray  # (d+1, d+1, nc)
    face_tang1: Optional[np.ndarray] = None  # (3,4,nc) 3-D only
    face_tang2: Optional[np.ndarray] = None
    cell_edges_len: Optional[np.ndarray] = None  # (3,nc) 2-D only
    face_area: Optional[np.ndarray] = None  # (d+1, nc)
    cell_procs: Optional[np.ndarray] = None
    face_boundary_index: Optional[np.ndarray] = None  # (d+1, nc) 1-based boundary element, -1 none
    nboundcells: int = 0
    boundary_area: Optional[np.ndarray] = None
    pg_names: Dict[str, int] = field(default_factory=dict)


def _neighbors_from_faces(cell_nodes0: np.ndarray, face_local: np.ndarray, nn: int) -> np.ndarray:
    """cell_nodes0: (nc, d+1) 0-based.  Returns (nc, nf) neighbor cell (0-based) or -1."""
    nc, nv = cell_nodes0.shape
    nf = face_local.shape[0]
    faces = np.sort(cell_nodes0[:, face_local], axis=2)  # (nc, nf, k)
    k = faces.shape[2]
    if float(nn) ** k < 9.0e18:
        key = np.zeros((nc, nf), dtype=np.int64)
        for a in range(k):
            key = key * nn + faces[:, :, a]
        key = key.ravel()
        order = np.argsort(key, kind="stable")
        ks = key[order]
        same = ks[1:] == ks[:-1]
    else:
        flat = faces.reshape(-1, k)
        order = np.lexsort(tuple(flat[:, a] for a in range(k - 1, -1, -1)))
        fs = flat[order]
        same = np.all(fs[1:] == fs[:-1], axis=1)
    nb = np.full(nc * nf, -1, dtype=np.int64)
    i = np.nonzero(same)[0]
    a, b = order[i], order[i + 1]
    nb[a] = b // nf
    nb[b] = a // nf
    return nb.reshape(nc, nf)


def _finish_boundary(m: Mesh):
    """Boundary sub-mesh numbering for the flux tallies: one element per boundary face, in (cell, face) order."""
    is_b = (m.cell_neighbors == -1) & (m.cell_faces_pg != -1)
    bidx = np.full(m.cell_neighbors.shape, -1, dtype=np.int32)
    flat = is_b.ravel(order="F")
    n = int(flat.sum())
    tmp = np.full(flat.shape, -1, dtype=np.int32)
    tmp[flat] = np.arange(1, n + 1, dtype=np.int32)
    bidx = tmp.reshape(m.cell_neighbors.shape, order="F")
    m.face_boundary_index = bidx
    m.nboundcells = n
    m.boundary_area = m.face_area.ravel(order="F")[flat].copy() if m.face_area is not None else np.ones(n)


def _tet_geometry(P: np.ndarray, fn, t1, t2, area, bc, vol):
    """Face normals/tangents/areas, volume and P1 basis coefficients of the tets P (n,4,3), written into the output
    slices.  Formulas: src/grid_and_partition.f90:1843-1850, 1932-1983, 1992-2039."""
    A = P[:, 1] - P[:, 0]
    B = P[:, 2] - P[:, 0]
    Cc = P[:, 3] - P[:, 0]
    det = Cc[:, 0] * (A[:, 1] * B[:, 2] - A[:, 2] * B[:, 1]) + Cc[:, 1] * (A[:, 2] * B[:, 0] - A[:, 0] * B[:, 2]) + \
        Cc[:, 2] * (A[:, 0] * B[:, 1] - A[:, 1] * B[:, 0])
    if np.any(det <= 0):
        raise ValueError("tets must be positively oriented (SURVEY App. E)")
    vol[:] = np.abs(det) / 6.0
    col = [[np.ascontiguousarray(P[:, v, k]) for k in range(3)] for v in range(4)]  # col[node][axis], contiguous
    for j in range(4):
        # explicit component formulas on contiguous columns; same operations, in the same order, as
        # np.cross / np.linalg.norm on the (n,3) arrays, so the values are bit-identical
        ia, ib, ic_ = (int(FACE_LOCAL_3D[j, i]) for i in range(3))
        u = [col[ib][k] - col[ia][k] for k in range(3)]
        w = [col[ic_][k] - col[ia][k] for k in range(3)]
        cr = [u[1] * w[2] - u[2] * w[1], u[2] * w[0] - u[0] * w[2], u[0] * w[1] - u[1] * w[0]]
        nrm = np.sqrt(cr[0] * cr[0] + cr[1] * cr[1] + cr[2] * cr[2])
        f = [cr[k] / nrm for k in range(3)]
        area[:, j] = 0.5 * nrm
        un = np.sqrt(u[0] * u[0] + u[1] * u[1] + u[2] * u[2])
        ta = [u[k] / un for k in range(3)]
        tb = [f[1] * ta[2] - f[2] * ta[1], f[2] * ta[0] - f[0] * ta[2], f[0] * ta[1] - f[1] * ta[0]]
        for k in range(3):
            fn[:, j, k] = f[k]
            t1[:, j, k] = ta[k]
            t2[:, j, k] = tb[k]
    (X1, Y1, Z1), (X2, Y2, Z2), (X3, Y3, Z3), (X4, Y4, Z4) = col
    bc[:, 0, 0] = Y2 * Z3 - Y3 * Z2 - Y2 * Z4 + Y4 * Z2 + Y3 * Z4 - Y4 * Z3
    bc[:, 0, 1] = -X2 * Z3 + X3 * Z2 + X2 * Z4 - X4 * Z2 - X3 * Z4 + X4 * Z3
    bc[:, 0, 2] = X2 * Y3 - X3 * Y2 - X2 * Y4 + X4 * Y2 + X3 * Y4 - X4 * Y3
    bc[:, 0, 3] = -X2 * Y3 * Z4 + X3 * Y2 * Z4 + X2 * Y4 * Z3 - X4 * Y2 * Z3 - X3 * Y4 * Z2 + X4 * Y3 * Z2
    bc[:, 1, 0] = -Y1 * Z3 + Y3 * Z1 + Y1 * Z4 - Y4 * Z1 - Y3 * Z4 + Y4 * Z3
    bc[:, 1, 1] = X1 * Z3 - X3 * Z1 - X1 * Z4 + X4 * Z1 + X3 * Z4 - X4 * Z3
    bc[:, 1, 2] = -X1 * Y3 + X3 * Y1 + X1 * Y4 - X4 * Y1 - X3 * Y4 + X4 * Y3
    bc[:, 1, 3] = X1 * Y3 * Z4 - X3 * Y1 * Z4 - X1 * Y4 * Z3 + X4 * Y1 * Z3 + X3 * Y4 * Z1 - X4 * Y3 * Z1
    bc[:, 2, 0] = Y1 * Z2 - Y2 * Z1 - Y1 * Z4 + Y4 * Z1 + Y2 * Z4 - Y4 * Z2
    bc[:, 2, 1] = -X1 * Z2 + X2 * Z1 + X1 * Z4 - X4 * Z1 - X2 * Z4 + X4 * Z2
    bc[:, 2, 2] = X1 * Y2 - X2 * Y1 - X1 * Y4 + X4 * Y1 + X2 * Y4 - X4 * Y2
    bc[:, 2, 3] = -X1 * Y2 * Z4 + X2 * Y1 * Z4 + X1 * Y4 * Z2 - X4 * Y1 * Z2 - X2 * Y4 * Z1 + X4 * Y2 * Z1
    bc[:, 3, 0] = -Y1 * Z2 + Y2 * Z1 + Y1 * Z3 - Y3 * Z1 - Y2 * Z3 + Y3 * Z2
    bc[:, 3, 1] = X1 * Z2 - X2 * Z1 - X1 * Z3 + X3 * Z1 + X2 * Z3 - X3 * Z2
    bc[:, 3, 2] = -X1 * Y2 + X2 * Y1 + X1 * Y3 - X3 * Y1 - X2 * Y3 + X3 * Y2
    bc[:, 3, 3] = X1 * Y2 * Z3 - X2 * Y1 * Z3 - X1 * Y3 * Z2 + X3 * Y1 * Z2 + X2 * Y3 * Z1 - X3 * Y2 * Z1
    bc[:] = -bc / 6.0 / vol[:, None, None]


def derive_3d(xyz: np.ndarray, tets0: np.ndarray, face_pg: np.ndarray, cell_pg: Optional[np.ndarray] = None) -> Mesh:
    """xyz (nn,3); tets0 (nc,4) 0-based, positively oriented; face_pg (nc,4) 1-based pg or -1."""
    nn, nc = xyz.shape[0], tets0.shape[0]
    nb = _neighbors_from_faces(tets0, FACE_LOCAL_3D, nn)
    fn = np.zeros((nc, 4, 3))
    t1 = np.zeros((nc, 4, 3))
    t2 = np.zeros((nc, 4, 3))
    area = np.zeros((nc, 4))
    bc = np.zeros((nc, 4, 4))  # [cell, basis i, coeff j]
    vol = np.zeros(nc)
    # in chunks: the ~450 element-wise operations per cell are memory-bound, and with cache-sized temporaries they
    # run ~4x faster than on 10^7-element arrays (element-wise, so the values do not depend on the chunking)
    CH = 1 << 14
    for a in range(0, nc, CH):
        b = min(nc, a + CH)
        _tet_geometry(xyz[tets0[a:b]], fn[a:b], t1[a:b], t2[a:b], area[a:b], bc[a:b], vol[a:b])
    nbr1 = np.where(nb >= 0, nb + 1, -1).astype(np.int32)
    m = Mesh(
        dims=3, nn=nn, nc=nc,
        node_coords=np.asfortranarray(xyz.T.astype(np.float64)),
        cell_nodes=np.asfortranarray((tets0 + 1).T.astype(np.int32)),
        cell_neighbors=np.asfortranarray(nbr1.T),
        face_normal=np.asfortranarray(fn.transpose(2, 1, 0)),
        face_tang1=np.asfortranarray(t1.transpose(2, 1, 0)),
        face_tang2=np.asfortranarray(t2.transpose(2, 1, 0)),
        cell_faces_pg=np.asfortranarray(face_pg.T.astype(np.int32)),
        cell_pg=(cell_pg if cell_pg is not None else np.full(nc, -1)).astype(np.int32),
        cell_volumes=vol,
        basis_coeffs=np.asfortranarray(bc.transpose(2, 1, 0)),
        face_area=np.asfortranarray(area.T),
    )
    _finish_boundary(m)
    return m


def box_tets(nx: int, ny: int, nz: int, lo=(0.0, 0.0, 0.0), hi=(1.0, 1.0, 1.0), jitter: float = 0.0,
             seed: int = 0) -> Mesh:
    """Structured-to-tet box (6 Kuhn tets per cube, conforming).  Physical groups:
    1 xmin, 2 xmax, 3 ymin, 4 ymax, 5 zmin, 6 zmax.  ``jitter`` perturbs interior nodes
    (fraction of the cell size) to make a genuinely irregular mesh."""
    gx = np.linspace(lo[0], hi[0], nx + 1)
    gy = np.linspace(lo[1], hi[1], ny + 1)
    gz = np.linspace(lo[2], hi[2], nz + 1)
    I, J, K = np.meshgrid(np.arange(nx + 1), np.arange(ny + 1), np.arange(nz + 1), indexing="ij")
    nid = (I + (nx + 1) * (J + (ny + 1) * K)).astype(np.int64)
    xyz = np.zeros(((nx + 1) * (ny + 1) * (nz + 1), 3))
    xyz[nid.ravel(), 0] = gx[I.ravel()]
    xyz[nid.ravel(), 1] = gy[J.ravel()]
    xyz[nid.ravel(), 2] = gz[K.ravel()]
    if jitter > 0:
        rng = np.random.default_rng(seed)
        h = np.array([(hi[0] - lo[0]) / nx, (hi[1] - lo[1]) / ny, (hi[2] - lo[2]) / nz])
        interior = ((I > 0) & (I < nx) & (J > 0) & (J < ny) & (K > 0) & (K < nz)).ravel()
        d = (rng.random((xyz.shape[0], 3)) - 0.5) * 2 * jitter * h
        xyz[nid.ravel()[interior]] += d[nid.ravel()[interior]]
    ci, cj, ck = np.meshgrid(np.arange(nx), np.arange(ny), np.arange(nz), indexing="ij")
    ci, cj, ck = ci.ravel(order="F"), cj.ravel(order="F"), ck.ravel(order="F")  # x fastest
    ncube = ci.size

    def node(di, dj, dk):
        return (ci + di) + (nx + 1) * ((cj + dj) + (ny + 1) * (ck + dk))

    perms = [(0, 1, 2), (0, 2, 1), (1, 0, 2), (1, 2, 0), (2, 0, 1), (2, 1, 0)]
    tets = np.zeros((ncube, 6, 4), dtype=np.int64)
    for t, p in enumerate(perms):
        d = np.zeros((4, 3), dtype=int)
        for s in range(3):
            d[s + 1] = d[s]
            d[s + 1, p[s]] += 1
        v = [node(*d[s]) for s in range(4)]
        # parity of the permutation decides the orientation; swap two nodes for odd ones
        inv = sum(1 for a in range(3) for b in range(a + 1, 3) if p[a] > p[b])
        if inv % 2 == 1:
            v[2], v[3] = v[3], v[2]
        tets[:, t, :] = np.stack(v, axis=1)
    tets0 = tets.reshape(-1, 4)
    # boundary tags from face-node coordinates (exact: grid planes are untouched by jitter)
    nc = tets0.shape[0]
    Ii = (tets0 % (nx + 1))
    Jj = (tets0 // (nx + 1)) % (ny + 1)
    Kk = tets0 // ((nx + 1) * (ny + 1))
    face_pg = np.full((nc, 4), -1, dtype=np.int32)
    for j in range(4):
        fl = FACE_LOCAL_3D[j]
        for pg, (arr, val) in enumerate([(Ii, 0), (Ii, nx), (Jj, 0), (Jj, ny), (Kk, 0), (Kk, nz)], start=1):
            on = np.all(arr[:, fl] == val, axis=1)
            face_pg[on, j] = pg
    m = derive_3d(xyz, tets0, face_pg)
    m.pg_names = {"xmin": 1, "xmax": 2, "ymin": 3, "ymax": 4, "zmin": 5, "zmax": 6}
    return m


def derive_2d(xy: np.ndarray, tris0: np.ndarray, edge_pg: np.ndarray, zmin: float, zmax: float, axi: bool = False,
              cell_pg: Optional[np.ndarray] = None) -> Mesh:
    nn, nc = xy.shape[0], tris0.shape[0]
    P = xy[tris0]  # (nc,3,2)
    x, y = P[:, :, 0], P[:, :, 1]
    cr = (x[:, 1] - x[:, 0]) * (y[:, 2] - y[:, 0]) - (x[:, 2] - x[:, 0]) * (y[:, 1] - y[:, 0])
    if np.any(cr < 0):
        raise ValueError("2-D triangles must be counter-clockwise (grid_and_partition.f90:1427-1439)")
    area = 0.5 * np.abs(x[:, 0] * (y[:, 1] - y[:, 2]) + x[:, 1] * (y[:, 2] - y[:, 0]) + x[:, 2] * (y[:, 0] - y[:, 1]))
    if axi:
        rad = (y[:, 0] + y[:, 1] + y[:, 2]) / 3.0
        vol = area * (zmax - zmin) * rad
    else:
        vol = area * (zmax - zmin)
    nb = _neighbors_from_faces(tris0, EDGE_LOCAL_2D, nn)
    en = np.zeros((nc, 3, 3))
    elen = np.zeros((nc, 3))
    for j in range(3):
        a, b = EDGE_LOCAL_2D[j]
        ln = np.sqrt((y[:, b] - y[:, a]) * (y[:, b] - y[:, a]) + (x[:, b] - x[:, a]) * (x[:, b] - x[:, a]))
        elen[:, j] = ln
        en[:, j, 0] = (y[:, b] - y[:, a]) / ln
        en[:, j, 1] = (x[:, a] - x[:, b]) / ln
    X1, X2, X3 = x[:, 0], x[:, 1], x[:, 2]
    Y1, Y2, Y3 = y[:, 0], y[:, 1], y[:, 2]
    bc = np.zeros((nc, 3, 3))
    bc[:, 0, 0] = Y2 - Y3
    bc[:, 0, 1] = -(X2 - X3)
    bc[:, 0, 2] = X2 * Y3 - X3 * Y2
    bc[:, 1, 0] = -(Y1 - Y3)
    bc[:, 1, 1] = X1 - X3
    bc[:, 1, 2] = X3 * Y1 - X1 * Y3
    bc[:, 2, 0] = -(Y2 - Y1)
    bc[:, 2, 1] = X2 - X1
    bc[:, 2, 2] = X1 * Y2 - X2 * Y1
    bc = 0.5 * bc / area[:, None, None]
    xyz = np.zeros((nn, 3))
    xyz[:, :2] = xy
    # boundary "area" of an edge: len*(zmax-zmin) [*rbar] (grid_and_partition.f90:1564-1569)
    farea = elen * (zmax - zmin)
    if axi:
        for j in range(3):
            a, b = EDGE_LOCAL_2D[j]
            farea[:, j] = elen[:, j] * (zmax - zmin) * 0.5 * (y[:, a] + y[:, b])
    m = Mesh(
        dims=2, nn=nn, nc=nc,
        node_coords=np.asfortranarray(xyz.T),
        cell_nodes=np.asfortranarray((tris0 + 1).T.astype(np.int32)),
        cell_neighbors=np.asfortranarray(np.where(nb >= 0, nb + 1, -1).astype(np.int32).T),
        face_normal=np.asfortranarray(en.transpose(2, 1, 0)),
        cell_edges_len=np.asfortranarray(elen.T),
        cell_faces_pg=np.asfortranarray(edge_pg.T.astype(np.int32)),
        cell_pg=(cell_pg if cell_pg is not None else np.full(nc, -1)).astype(np.int32),
        cell_volumes=vol,
        basis_coeffs=np.asfortranarray(bc.transpose(2, 1, 0)),
        face_area=np.asfortranarray(farea.T),
    )
    _finish_boundary(m)
    return m


def mapped_quads_to_tris(X: np.ndarray, Y: np.ndarray, zmin=-0.5, zmax=0.5, axi=False, periodic_j=False) -> Mesh:
    """Triangulate a logically rectangular node array X,Y of shape (ni+1, nj+1) (2 CCW triangles per quad).
    Physical groups: 1 = i-min side, 2 = i-max, 3 = j-min, 4 = j-max."""
    ni, nj = X.shape[0] - 1, X.shape[1] - 1
    nid = np.arange((ni + 1) * (nj + 1)).reshape(ni + 1, nj + 1)
    xy = np.stack([X.ravel(), Y.ravel()], axis=1)
    a = nid[:-1, :-1].ravel()
    b = nid[1:, :-1].ravel()
    c = nid[1:, 1:].ravel()
    d = nid[:-1, 1:].ravel()
    t1 = np.stack([a, b, c], axis=1)
    t2 = np.stack([a, c, d], axis=1)
    tris0 = np.empty((2 * a.size, 3), dtype=np.int64)
    tris0[0::2] = t1
    tris0[1::2] = t2
    # orientation: flip if the mapping is left-handed
    P = xy[tris0]
    cr = (P[:, 1, 0] - P[:, 0, 0]) * (P[:, 2, 1] - P[:, 0, 1]) - (P[:, 2, 0] - P[:, 0, 0]) * (P[:, 1, 1] - P[:, 0, 1])
    flip = cr < 0
    tris0[flip] = tris0[flip][:, [0, 2, 1]]
    ii = tris0 // (nj + 1)
    jj = tris0 % (nj + 1)
    edge_pg = np.full((tris0.shape[0], 3), -1, dtype=np.int32)
    for j in range(3):
        e = EDGE_LOCAL_2D[j]
        for pg, (arr, val) in enumerate([(ii, 0), (ii, ni), (jj, 0), (jj, nj)], start=1):
            on = np.all(arr[:, e] == val, axis=1)
            edge_pg[on, j] = pg
    m = derive_2d(xy, tris0, edge_pg, zmin, zmax, axi)
    m.pg_names = {"imin": 1, "imax": 2, "jmin": 3, "jmax": 4}
    return m


def rect_tris(nx: int, ny: int, lo=(0.0, 0.0), hi=(1.0, 1.0), zmin=-0.5, zmax=0.5, axi=False) -> Mesh:
    gx = np.linspace(lo[0], hi[0], nx + 1)
    gy = np.linspace(lo[1], hi[1], ny + 1)
    X, Y = np.meshgrid(gx, gy, indexing="ij")
    m = mapped_quads_to_tris(X, Y, zmin, zmax, axi)
    m.pg_names = {"xmin": 1, "xmax": 2, "ymin": 3, "ymax": 4}
    return m


def line_1d(nx: int, xlo: float, xhi: float, ymin=0.0, ymax=1.0, zmin=0.0, zmax=1.0) -> Mesh:
    """1-D SU2 line mesh.  Physical groups: 1 = left vertex, 2 = right vertex."""
    gx = np.linspace(xlo, xhi, nx + 1)
    nn, nc = nx + 1, nx
    xyz = np.zeros((nn, 3))
    xyz[:, 0] = gx
    cells0 = np.stack([np.arange(nx), np.arange(1, nx + 1)], axis=1)
    seglen = np.abs(gx[:-1] - gx[1:])
    vol = seglen * (ymax - ymin) * (zmax - zmin)
    nb = np.full((nc, 2), -1, dtype=np.int32)
    nb[1:, 0] = np.arange(1, nc)  # neighbour across node 1 = previous cell (1-based ids)
    nb[:-1, 1] = np.arange(2, nc + 1)
    en = np.zeros((nc, 2, 3))
    en[:, 0, 0] = -1.0
    en[:, 1, 0] = 1.0
    bc = np.zeros((nc, 2, 2))
    bc[:, 0, 0] = -1.0
    bc[:, 0, 1] = gx[1:]
    bc[:, 1, 0] = 1.0
    bc[:, 1, 1] = -gx[:-1]
    bc = bc / seglen[:, None, None]
    fpg = np.full((nc, 2), -1, dtype=np.int32)
    fpg[0, 0] = 1
    fpg[-1, 1] = 2
    area = np.full((nc, 2), (ymax - ymin) * (zmax - zmin))
    m = Mesh(
        dims=1, nn=nn, nc=nc,
        node_coords=np.asfortranarray(xyz.T),
        cell_nodes=np.asfortranarray((cells0 + 1).T.astype(np.int32)),
        cell_neighbors=np.asfortranarray(nb.T),
        face_normal=np.asfortranarray(en.transpose(2, 1, 0)),
        cell_faces_pg=np.asfortranarray(fpg.T),
        cell_pg=np.full(nc, -1, dtype=np.int32),
        cell_volumes=vol,
        basis_coeffs=np.asfortranarray(bc.transpose(2, 1, 0)),
        face_area=np.asfortranarray(area.T),
    )
    _finish_boundary(m)
    m.pg_names = {"left": 1, "right": 2}
    return m


def assign_cells_to_procs(m: Mesh, nranks: int, style: int = 0) -> np.ndarray:
    """ASSIGN_CELLS_TO_PROCS without load balance (src/grid_and_partition.f90:150-210):
    uniform strips of the cell-centroid coordinate (0 STRIPSX, 1 STRIPSY, 2 STRIPSZ)."""
    if nranks == 1:
        return np.zeros(m.nc, dtype=np.int32)
    cn = np.asarray(m.cell_nodes) - 1
    coords = np.asarray(m.node_coords)
    nv = m.dims + 1
    cen = coords[style, cn[0]]
    for i in range(1, nv):
        cen = cen + coords[style, cn[i]]
    cen = cen / float(nv)
    cmax, cmin = cen.max(), cen.min()
    # REAL(N_MPI_THREADS) is single precision in the reference; exact for small integers
    procs = ((cen - cmin) / (cmax - cmin) * float(np.float32(nranks))).astype(np.int64)
    procs = np.clip(procs, 0, nranks - 1)
    return procs.astype(np.int32)


def load_balance_cut(m: Mesh, np_cells: np.ndarray, nranks: int, style: int = 0) -> np.ndarray:
    """The LOAD_BALANCE branch of ASSIGN_CELLS_TO_PROCS (src/grid_and_partition.f90:214-310): cell weight = particles in
    the cell (summed over the ranks) + 10 % of the mean, cells ordered by centroid coordinate, greedy cut whenever the
    cumulated weight exceeds the per-rank share.  Host work in the reference (it stays in Fortran); the device side is
    pgpu_get_cell_counts -> (MPI_REDUCE on the host) -> pgpu_set_cell_procs -> EXCHANGE."""
    cn = np.asarray(m.cell_nodes) - 1
    coords = np.asarray(m.node_coords)
    nv = m.dims + 1
    cen = coords[style, cn[0]]
    for i in range(1, nv):
        cen = cen + coords[style, cn[i]]
    cen = cen / float(nv)
    weight = np.asarray(np_cells, dtype=np.float64)
    weight = weight + float(np.float32(0.1)) * weight.sum() / float(m.nc)
    per_proc = weight.sum() / float(nranks)
    order = np.argsort(cen, kind="stable")      # QUICKSORT by centroid (ties: any order; the regular test meshes have none across a cut)
    procs = np.zeros(m.nc, dtype=np.int32)
    iproc, cum = 0, 0.0
    for ic in order:
        cum += weight[ic]
        procs[ic] = min(iproc, nranks - 1)
        if cum > per_proc:
            iproc += 1
            cum = 0.0
    return procs


def to_su2(m: Mesh, path: str):
    """Write the mesh in the SU2 subset the reference readers accept."""
    etype = {1: 3, 2: 5, 3: 10}[m.dims]
    with open(path, "w") as f:
        f.write(f"NDIME= {m.dims}\n")
        f.write(f"NPOIN= {m.nn}\n")
        xyz = np.asarray(m.node_coords).T
        for p in xyz:
            f.write(" ".join(repr(float(v)) for v in p[: m.dims]) + "\n")
        f.write(f"NELEM= {m.nc}\n")
        cn = np.asarray(m.cell_nodes).T - 1
        for c in cn:
            f.write(f"{etype} " + " ".join(str(int(v)) for v in c) + "\n")
        pgs = sorted(set(int(v) for v in np.unique(m.cell_faces_pg) if v > 0))
        names = {v: k for k, v in m.pg_names.items()}
        f.write(f"NMARK= {len(pgs)}\n")
        fl = {3: FACE_LOCAL_3D, 2: EDGE_LOCAL_2D, 1: np.array([[0], [1]])}[m.dims]
        btype = {3: 5, 2: 3, 1: 1}[m.dims]
        fpg = np.asarray(m.cell_faces_pg).T
        for pg in pgs:
            sel = np.argwhere(fpg == pg)
            f.write(f"MARKER_TAG= {names.get(pg, 'pg%d' % pg)}\n")
            f.write(f"MARKER_ELEMS= {len(sel)}\n")
            for c, j in sel:
                f.write(f"{btype} " + " ".join(str(int(cn[c, k])) for k in fl[j]) + "\n")
