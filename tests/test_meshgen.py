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


def _fortran_quicksort(array, order, lo, hi):
    """QUICKSORT, src/tools.f90:1314-1339 (Lomuto partition, 1-based bounds, sorts `array` and carries `order`)."""
    if lo >= 1 and hi >= 1 and lo < hi:
        pivot = array[hi - 1]
        i = lo - 1
        for j in range(lo, hi):
            if array[j - 1] <= pivot:
                i += 1
                array[i - 1], array[j - 1] = array[j - 1], array[i - 1]
                order[i - 1], order[j - 1] = order[j - 1], order[i - 1]
        array[i], array[hi - 1] = array[hi - 1], array[i]
        order[i], order[hi - 1] = order[hi - 1], order[i]
        p = i + 1
        _fortran_quicksort(array, order, lo, p - 1)
        _fortran_quicksort(array, order, p + 1, hi)


def test_partition_helpers_match_a_plain_reading_of_assign_cells_to_procs():
    """ASSIGN_CELLS_TO_PROCS (src/grid_and_partition.f90:123-310) in plain Python, both branches, against the helpers the
    multi-GPU tests and the bench use: uniform strips of the centroid coordinate, and the particle-weighted re-cut (weights =
    particles per cell + 10 % of the mean, cells ordered by the reference's own QUICKSORT, greedy cut).  Jittered mesh: no two
    centroids coincide, so the order of ties (which the reference leaves to its unstable sort) plays no role."""
    import numpy as np
    from pantera_pic_dsmc_b200 import meshgen as mg
    m = mg.box_tets(6, 3, 2, lo=(0.0, 0.0, 0.0), hi=(3.0, 1.0, 0.7), jitter=0.2, seed=5)
    xyz = np.asarray(m.node_coords).copy()
    xyz += np.random.default_rng(1).random(xyz.shape) * 1e-3      # every node moved a little (only the centroids matter here)
    m.node_coords = np.asfortranarray(xyz)
    cn = np.asarray(m.cell_nodes) - 1
    nr = 4
    for style in (0, 1, 2):
        cen = [float((xyz[style, cn[0, i]] + xyz[style, cn[1, i]] + xyz[style, cn[2, i]] + xyz[style, cn[3, i]]) / 4.0) for i in range(m.nc)]
        cmax, cmin = max(cen), min(cen)
        want = []
        for i in range(m.nc):
            p = int((cen[i] - cmin) / (cmax - cmin) * float(np.float32(nr)))
            want.append(min(max(p, 0), nr - 1))
        assert list(mg.assign_cells_to_procs(m, nr, style)) == want
        assert len(set(cen)) == m.nc
        rng = np.random.default_rng(style)
        np_cells = rng.poisson(30.0 * (1.0 + 3.0 * (np.array(cen) - cmin) / (cmax - cmin)))   # more particles downstream
        weight = [float(v) for v in np_cells]
        tot = sum(weight)
        weight = [v + float(np.float32(0.1)) * tot / float(m.nc) for v in weight]
        per_proc = sum(weight) / float(nr)
        arr, order = list(cen), list(range(1, m.nc + 1))
        _fortran_quicksort(arr, order, 1, m.nc)
        assert arr == sorted(cen)
        want, iproc, cum = [0] * m.nc, 0, 0.0
        for i in range(m.nc):
            ic = order[i]
            cum = cum + weight[ic - 1]
            want[ic - 1] = iproc
            if cum > per_proc:
                iproc += 1
                cum = 0.0
        got = mg.load_balance_cut(m, np_cells, nr, style)
        assert max(want) == nr - 1 and list(got) == want
        loads = [float(np_cells[np.array(want) == r].sum()) for r in range(nr)]
        assert max(loads) < 1.35 * (sum(loads) / nr)                                           # the re-cut balances the particles
