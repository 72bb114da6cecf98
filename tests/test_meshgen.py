"""Invariants of the tetrahedral mesh derivation (the host-side stand-in for READ_3D_UNSTRUCTURED_GRID_SU2's geometry,
src/grid_and_partition.f90:1843-2039) on a jittered box large enough to span several evaluation chunks."""
import numpy as np

from pantera_pic_dsmc_b200 import meshgen as mg


def test_box_tets_geometry_invariants():
    nx, ny, nz = 30, 28, 26                       # 131040 tets > 2^17: eight chunks of 2^14
    m = mg.box_tets(nx, ny, nz, lo=(0.0, -1.0, 2.0), hi=(3.0, 1.0, 4.5), jitter=0.2, seed=5)
    assert m.nc == 6 * nx * ny * nz
    fn, area = np.asarray(m.face_normal), np.asarray(m.face_area)
    # closed surface of every tet: sum_f A_f n_f = 0; unit normals; volumes tile the box
    closure = np.einsum("kfc,fc->kc", fn, area)
    assert np.abs(closure).max() < 1e-12 * area.max()
    assert np.allclose(np.sum(fn * fn, axis=0), 1.0, rtol=0, atol=1e-14)
    assert np.isclose(m.cell_volumes.sum(), 3.0 * 2.0 * 2.5, rtol=1e-12)
    # outward: the opposite vertex lies on the negative side of every face (face j is opposite... of local node map)
    xyz, cn = np.asarray(m.node_coords), np.asarray(m.cell_nodes) - 1
    cen = xyz[:, cn].mean(axis=1)                                # (3, nc) centroids
    for j in range(4):
        pj = xyz[:, cn[j]]                                       # reference point of face j = local node j (ADVECT's P_i)
        assert np.all(np.sum((pj - cen) * fn[:, j], axis=0) > 0)  # the centroid is inside: DX > 0 on every face
    # neighbours are symmetric and boundary faces carry their physical group
    nb = np.asarray(m.cell_neighbors)
    c = np.arange(m.nc)
    for j in range(4):
        has = nb[j] > 0
        back = nb[:, nb[j, has] - 1]
        assert np.all(np.any(back == (c[has] + 1)[None, :], axis=0))
    assert np.all((nb == -1) == (np.asarray(m.cell_faces_pg) != -1))
    # tangents: orthonormal frame with the normal
    t1, t2 = np.asarray(m.face_tang1), np.asarray(m.face_tang2)
    assert np.abs(np.sum(t1 * fn, axis=0)).max() < 1e-13 and np.abs(np.sum(t2 * fn, axis=0)).max() < 1e-13
    assert np.abs(np.sum(t1 * t2, axis=0)).max() < 1e-13
    # P1 basis: psi_p(x_q) = delta_pq at the nodes
    bc = np.asarray(m.basis_coeffs)                               # (coeff, basis p, cell)
    for q in range(4):
        xq = xyz[:, cn[q]]
        psi = bc[0] * xq[0] + bc[1] * xq[1] + bc[2] * xq[2] + bc[3]
        want = np.zeros_like(psi)
        want[q] = 1.0
        assert np.abs(psi - want).max() < 1e-9
