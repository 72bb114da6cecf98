"""EXCHANGE parity against the oracle's rank-per-strip model.

The reference runs one MPI rank per strip: ADVECT, then EXCHANGE ships every particle whose cell belongs to another
rank (src/grid_and_partition.f90:525-622), then the collisions.  oracle/po_mpi_sim.c is that model on host threads.

Single GPU (what the driver's box has): NRANKS library contexts on ONE device, the all-to-all carried by the host
through pgpu_exchange_pack / pgpu_append_particles -- the split a Fortran host that keeps its own MPI_ALLTOALLV
would use.  This exercises everything of EXCHANGE except the NCCL transport: ownership lookup, pack by destination,
tombstoning, append, and the sort that follows.
Two GPUs (run with `gpurun --gpus 2` by the builder; skipped on smaller boxes): the same comparisons through
pgpu_exchange / pgpu_run_steps over NCCL.

Compared particle for particle, matched by ID: with NO_COLL and specular walls everything is deterministic and must
be BIT-IDENTICAL after several steps (positions, velocities, cells, and which rank holds which particle); with
collisions each rank's Philox streams are seeded RNG_SEED_GLOBAL*(PROC_ID+1) (src/initialization.f90:2294-2296) and
the oracle collides the device's sorted order, value for value.
"""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from pantera_pic_dsmc_b200 import capi, workloads as wl  # noqa: E402
from oracle import refbench  # noqa: E402
from oracle.binding import load_oracle  # noqa: E402

import util  # noqa: E402

pytestmark = pytest.mark.gpu

CUBES, PPC = (4, 5, 5), 25
EXACT = ("X", "Y", "Z", "VX", "VY", "VZ", "EROT", "EVIB")


def strip_workload(rank, world, collisions, mesh=None):
    # The per-step NCCL EXCHANGE only talks to ranks that share a face with this one (pgpu_exchange.cu): with four
    # strips the x extent is doubled and the step shortened, so that no particle can cross a whole strip in one step
    # (which would fail the step loudly, as it should).
    cubes, dt_frac = (CUBES, 0.6) if world < 4 else ((2 * CUBES[0], CUBES[1] - 1, CUBES[2] - 1), 0.3)
    n = cubes[0] * cubes[1] * cubes[2] * 6 * PPC
    w = wl.c5_channel_strip(*cubes, n=n, rank=rank, world=world, seed=31 + rank, mesh=mesh, dt_frac=dt_frac)
    w.cfg.track_id = 1
    w.cfg.capacity = 3 * n
    if collisions:
        w.cfg.collision_type = capi.DSMC
    else:
        w.cfg.collision_type = capi.NO_COLL
        w.bcs["particle_bc"] = [capi.SPECULAR] * 6
    return w


def host_all_to_all(libs):
    """MPI_ALLTOALL + MPI_ALLTOALLV by hand: rank r appends what every other rank packed for it, sources ascending."""
    world = len(libs)
    packed = [lib.exchange_pack(world) for lib in libs]
    moved = 0
    for dst in range(world):
        for src in range(world):
            if src == dst:
                continue
            buf, cnt = packed[src]
            off = int(cnt[:dst].sum())
            n = int(cnt[dst])
            if n:
                libs[dst].append_particles(np.ascontiguousarray(buf[off:off + n]))
                moved += n
    return moved


def compare_ranks(dev, orc, exact):
    for r, (a, b) in enumerate(zip(orc, dev)):
        a, b = util.by_id(a), util.by_id(b)
        assert np.array_equal(a["ID"], b["ID"]), f"rank {r}: the device and the oracle hold different particles"
        if exact:
            util.assert_particles_match(a, b, exact_fields=EXACT, what=f"rank {r}")
        else:
            util.assert_particles_match(a, b, rtol=1e-12, atol=1e-9, what=f"rank {r}")


@pytest.mark.parametrize("world", [2, 3])
def test_one_gpu_contexts_match_rank_per_strip_oracle_bit_exact(world):
    steps = 4
    gpus, orcs, mesh = [], [], None
    try:
        for r in range(world):
            w = strip_workload(r, world, collisions=False, mesh=mesh)
            mesh = w.mesh
            g, o = capi.load_gpu(0), load_oracle(rng_mode=0)
            wl.apply(g, w)
            wl.apply(o, w)
            gpus.append(g)
            orcs.append(o)
        moved = 0
        for s in range(steps):
            for g in gpus:
                g.advect(1 + s)
            moved += host_all_to_all(gpus)
        sec, tot = refbench.sim_run(orcs, 1, steps)            # oracle/po_mpi_sim.c, NO_COLL
        assert moved == tot[2] and moved > 100, (moved, tot)
        dev = [g.download_particles() for g in gpus]
        ref = [o.download_particles() for o in orcs]
        compare_ranks(dev, ref, exact=True)
        owner = np.asarray(mesh.cell_procs)
        for r, p in enumerate(dev):
            assert np.all(owner[p["IC"] - 1] == r), f"rank {r} holds particles of cells it does not own"
        assert sum(len(p) for p in dev) == world * len(w.particles)
    finally:
        for lib in gpus + orcs:
            lib.close()


def test_one_gpu_contexts_collisions_per_rank_streams():
    """ADVECT -> EXCHANGE -> sort -> DSMC on two rank contexts; every rank's collisions value for value against the
    oracle on that rank's Philox streams (seed * (rank + 1)), three steps, bench wall models (diffuse in y)."""
    world, steps = 2, 3
    gpus, orcs, mesh = [], [], None
    try:
        for r in range(world):
            w = strip_workload(r, world, collisions=True, mesh=mesh)
            mesh = w.mesh
            g, o = capi.load_gpu(0), load_oracle(rng_mode=1)
            wl.apply(g, w)
            wl.apply(o, w, upload=False)
            gpus.append(g)
            orcs.append(o)
        ncoll = 0
        for s in range(steps):
            t = 1 + s
            for g in gpus:
                g.advect(t)
            assert host_all_to_all(gpus) > 0
            for r, (g, o) in enumerate(zip(gpus, orcs)):
                g.sort()
                ps = g.download_particles()              # the device's sorted order is the order the pairs index into
                o.upload_particles(ps)
                co, cg = o.collide(t), g.collide(t)
                assert co == cg and cg[0] > 0, (r, t, co, cg)
                ncoll += cg[0]
                a, b = o.download_particles(), g.download_particles()
                assert np.array_equal(a["ID"], b["ID"])
                util.assert_particles_match(a, b, rtol=1e-12, atol=1e-9, what=f"rank {r} step {t}")
        assert ncoll > 100
    finally:
        for lib in gpus + orcs:
            lib.close()


def test_load_balance_recut_then_exchange():
    """LOAD_BALANCE (src/grid_and_partition.f90:214-310, called from src/timecycle.f90:185-189): per-cell counts from the
    device, summed over the ranks on the host, the reference's greedy cut, pgpu_set_cell_procs, a full EXCHANGE.  The
    drift of the channel has piled particles up downstream; after the re-cut the ranks are balanced again, every
    particle sits on its new owner, and the oracle's rank model lands on the same particles bit for bit."""
    from pantera_pic_dsmc_b200 import meshgen as mg
    world, steps = 2, 6
    gpus, orcs, mesh = [], [], None
    try:
        for r in range(world):
            w = strip_workload(r, world, collisions=False, mesh=mesh)
            mesh = w.mesh
            g, o = capi.load_gpu(0), load_oracle(rng_mode=0)
            wl.apply(g, w)
            wl.apply(o, w)
            gpus.append(g)
            orcs.append(o)
        for s in range(steps):
            for g in gpus:
                g.advect(1 + s)
            host_all_to_all(gpus)
        refbench.sim_run(orcs, 1, steps)
        before = [g.num_particles() for g in gpus]
        assert max(before) > 1.08 * min(before), before           # the drift has unbalanced the strips
        npc = np.zeros(mesh.nc, dtype=np.int64)
        for g in gpus:
            g.sort()
            npc += g.get_cell_counts(mesh.nc)
        npc_o = np.zeros(mesh.nc, dtype=np.int64)
        for o in orcs:
            p = o.download_particles()
            npc_o += np.bincount(p["IC"] - 1, minlength=mesh.nc)
        assert np.array_equal(npc, npc_o)
        procs = mg.load_balance_cut(mesh, npc, world, 0)
        assert set(np.unique(procs)) == {0, 1} and not np.array_equal(procs, np.asarray(mesh.cell_procs))
        for lib in gpus + orcs:
            lib.set_cell_procs(procs)
        host_all_to_all(gpus)
        host_all_to_all(orcs)
        after = [g.num_particles() for g in gpus]
        assert sum(after) == sum(before) and max(after) < 1.03 * min(after), (before, after)
        dev = [g.download_particles() for g in gpus]
        ref = [o.download_particles() for o in orcs]
        compare_ranks(dev, ref, exact=True)
        for r, p in enumerate(dev):
            assert np.all(procs[p["IC"] - 1] == r)
        # and the run goes on under the new ownership (hop codes rebuilt with the new flags)
        for s in range(2):
            for g in gpus:
                g.advect(50 + s)
            host_all_to_all(gpus)
        refbench.sim_run(orcs, 50, 2)
        compare_ranks([g.download_particles() for g in gpus], [o.download_particles() for o in orcs], exact=True)
    finally:
        for lib in gpus + orcs:
            lib.close()


# --------------------------------------------------------------------------------------------- NCCL (>= 2 GPUs)
def _free_port() -> int:
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _nccl_worker(rank, world, port, out_dir, p2p):
    import torch
    import torch.distributed as dist

    # "1": EXCHANGE over peer-memory windows (CUDA IPC, the default); "0": every message through ncclSend/ncclRecv
    # (the path a pair takes when its windows cannot be mapped)
    os.environ["PANTERA_GPU_XCHG_P2P"] = p2p
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    steps = 4
    # (a) NO_COLL, specular: pgpu_run_steps with the NCCL exchange inside
    w = strip_workload(rank, world, collisions=False)
    g = capi.load_gpu(rank)
    wl.apply(g, w)
    uid = [g.comm_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(uid, src=0)
    g.comm_init(world, rank, uid[0])
    g.run_steps(1, steps)
    np.save(os.path.join(out_dir, f"nocoll{rank}.npy"), g.download_particles())
    # (b) collisions, step by step: advect -> exchange (NCCL) -> sort -> collide, against this rank's oracle
    w2 = strip_workload(rank, world, collisions=True, mesh=w.mesh)
    wl.apply(g, w2)
    o = load_oracle(rng_mode=1)
    wl.apply(o, w2, upload=False)
    ok = True
    for s in range(3):
        t = 10 + s
        g.advect(t)
        g.exchange()
        g.sort()
        ps = g.download_particles()
        o.upload_particles(ps)
        co, cg = o.collide(t), g.collide(t)
        a, b = o.download_particles(), g.download_particles()
        ok = ok and co == cg and cg[0] > 0 and np.array_equal(a["ID"], b["ID"])
        for f in ("VX", "VY", "VZ"):
            ok = ok and bool(np.allclose(a[f], b[f], rtol=1e-12, atol=1e-9))
    open(os.path.join(out_dir, f"coll{rank}"), "w").write("1" if ok else "0")
    g.close()
    o.close()
    dist.destroy_process_group()


@pytest.mark.timeout(600)
@pytest.mark.parametrize("world,p2p", [(2, "1"), (2, "0"), (4, "1")])
def test_nccl_ranks_match_rank_per_strip_oracle(tmp_path, world, p2p):
    import torch
    import torch.multiprocessing as mp

    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    mp.spawn(_nccl_worker, args=(world, _free_port(), str(tmp_path), p2p), nprocs=world, join=True)
    orcs, mesh = [], None
    for r in range(world):
        w = strip_workload(r, world, collisions=False, mesh=mesh)
        mesh = w.mesh
        o = load_oracle(rng_mode=0)
        wl.apply(o, w)
        orcs.append(o)
    refbench.sim_run(orcs, 1, 4)
    ref = [o.download_particles() for o in orcs]
    dev = [np.load(tmp_path / f"nocoll{r}.npy") for r in range(world)]
    compare_ranks(dev, ref, exact=True)
    assert all(open(tmp_path / f"coll{r}").read() == "1" for r in range(world))
    for o in orcs:
        o.close()
