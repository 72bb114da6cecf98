"""The bench workloads that inject on the device every step (`bench.py --workload c2-cyl | c3 | c5-open`): their meshes are
valid, their EMIT tasks carry the free-stream flux INITINJECTION would compute (src/initialization.f90:2575-2675), and the CPU
leg of the bench (the oracle driven through `bench.alt_step_fn`) runs them with a population that stays put."""
import numpy as np
import pytest

import bench
from pantera_pic_dsmc_b200 import workloads as wl
from oracle.binding import load_oracle

KB = 1.380649e-23


def check_mesh(m):
    nb = np.asarray(m.cell_neighbors)
    assert np.all(np.asarray(m.cell_volumes) > 0)
    for j in range(m.dims + 1):                     # the neighbour relation is symmetric
        c = np.nonzero(nb[j] > 0)[0]
        back = nb[:, nb[j, c] - 1]
        assert np.all((back == (c + 1)[None, :]).any(axis=0))
    fpg = np.asarray(m.cell_faces_pg)
    assert np.all((nb == -1) == (fpg != -1))         # every boundary face carries a physical group, no interior face does


@pytest.mark.parametrize("name,n", [("c2-cyl", 60_000), ("c3", 60_000), ("c5-open", 61_440)])
def test_emitting_bench_workloads_run_on_the_cpu_leg(name, n):
    w = bench.build_alt(name, n, small=True)
    check_mesh(w.mesh)
    assert len(w.emit_tasks) > 0 and all(len(t["nfs"]) == len(w.mixtures[0]) for t in w.emit_tasks)
    # expected inflow per step: sum over the emitting faces of the Maxwellian one-sided flux x area x DT / FNUM
    per_step = sum(sum(t["nfs"]) for t in w.emit_tasks)
    assert per_step > 1.0
    o = load_oracle(rng_mode=0)
    try:
        wl.apply(o, w)
        wl.apply_sources(o, w)
        step = bench.alt_step_fn(name, o, w)
        n0 = o.num_particles()
        o.boundaries_emit(1)
        made = o.num_particles() - n0
        assert abs(made - per_step) < 6.0 * np.sqrt(per_step) + 2.0 * len(w.emit_tasks) ** 0.5 + 3.0
        ncoll, grp = step(2, 6)
        assert set(grp) == {"emit", "move", "sort", "collide"} and ncoll > 0
        n1 = o.num_particles()
        assert 0.85 * n0 < n1 < 1.25 * n0                # inflow and outflow roughly balance over a few steps
        p = o.download_particles()
        assert p["IC"].min() >= 1 and p["IC"].max() <= w.mesh.nc
    finally:
        o.close()


def test_cylinder_free_stream_flux_matches_the_geometry():
    """Sum of NFS over the outer boundary = n_inf u_inf x (projected height of the half annulus) x depth x DT / FNUM up to
    the thermal back-flux (speed ratio 20: negligible) -- an independent check of emit_tasks_for_group on a curved boundary."""
    w = wl.c2_cylinder(n=50_000, ni=30, nj=120, with_reactions=False)
    nrho, u = 1e20, 7000.0
    total = sum(sum(t["nfs"]) for t in w.emit_tasks)
    expect = nrho * u * (0.4572 * 1.0) * w.cfg.dt / w.cfg.fnum      # the free stream crosses the projected height r_out (upstream half)
    assert total == pytest.approx(expect, rel=2e-3)
