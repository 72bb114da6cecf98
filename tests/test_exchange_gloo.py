"""N > 1 host logic on CPU: two processes (torch.distributed, gloo, world size 2), one x-strip each.

Every rank drives an oracle context through the reference's per-step sequence ADVECT -> EXCHANGE -> collisions
(src/timecycle.f90:271-299); EXCHANGE = pack by destination (src/grid_and_partition.f90:562-583) + all-to-all of
counts and records (:590-605) + append (:610-612), the all-to-all carried by gloo.  Checked: particles are
conserved, every particle ends on the rank that owns its cell, and the distributed run equals the in-process
thread model of oracle/po_mpi_sim.c (same per-rank MT19937-64 seeds) particle for particle.
"""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from pantera_pic_dsmc_b200 import capi, workloads as wl  # noqa: E402
from oracle import refbench  # noqa: E402
from oracle.binding import load_oracle  # noqa: E402

CUBES, PPC, STEPS = (3, 4, 4), 12, 4


def _free_port() -> int:
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _exchange(o, rank, world):
    """EXCHANGE through gloo: returns the number of particles received."""
    buf, cnt = o.exchange_pack(world)
    counts = torch.from_numpy(cnt.astype(np.int64))
    allc = [torch.zeros(world, dtype=torch.int64) for _ in range(world)]
    dist.all_gather(allc, counts)                      # MPI_ALLTOALL of the counts
    raw = torch.from_numpy(np.frombuffer(buf.tobytes(), dtype=np.uint8).copy())
    rec = buf.dtype.itemsize
    offs = np.concatenate([[0], np.cumsum(cnt)])
    got = 0
    for src in range(world):                           # MPI_ALLTOALLV as point-to-point pairs
        for dst in range(world):
            if src == dst:
                continue
            n = int(allc[src][dst])
            if n == 0:
                continue
            if rank == src:
                dist.send(raw[offs[dst] * rec:(offs[dst] + n) * rec].contiguous(), dst=dst)
            elif rank == dst:
                t = torch.empty(n * rec, dtype=torch.uint8)
                dist.recv(t, src=src)
                o.append_particles(np.frombuffer(t.numpy().tobytes(), dtype=buf.dtype))
                got += n
    return got


def _worker(rank, world, port, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    n = CUBES[0] * CUBES[1] * CUBES[2] * 6 * PPC
    w = wl.c5_channel_strip(*CUBES, n=n, rank=rank, world=world, seed=7 + rank)
    w.cfg.track_id = 1
    o = load_oracle(rng_mode=0)
    wl.apply(o, w)
    moved = 0
    for s in range(STEPS):
        o.advect(1 + s)
        moved += _exchange(o, rank, world)
        o.collide(1 + s)
    p = o.download_particles()
    tot = torch.tensor([len(p), moved], dtype=torch.int64)
    dist.all_reduce(tot)
    owner = np.asarray(w.mesh.cell_procs)[p["IC"] - 1]
    np.save(os.path.join(out_dir, f"rank{rank}.npy"), p)
    ok = bool(np.all(owner == rank)) and int(tot[0]) == world * n and int(tot[1]) > 0
    with open(os.path.join(out_dir, f"ok{rank}"), "w") as f:
        f.write("1" if ok else f"0 owner_ok={np.all(owner == rank)} total={int(tot[0])} moved={int(tot[1])}")
    o.close()
    dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_two_rank_gloo_exchange_matches_thread_model(tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    for r in range(world):
        assert open(tmp_path / f"ok{r}").read() == "1"
    # the same run in the in-process thread model of the reference's MPI execution (oracle/po_mpi_sim.c)
    libs, n, _ = refbench.make_ranks(world, CUBES, PPC, seed=7)
    for lib in libs:
        pass
    refbench.sim_run(libs, 1, STEPS)
    for r, lib in enumerate(libs):
        a = lib.download_particles()
        b = np.load(tmp_path / f"rank{r}.npy")
        assert len(a) == len(b)
        for f in ("X", "Y", "Z", "VX", "VY", "VZ", "IC", "S_ID"):
            assert np.array_equal(np.sort(a[f]), np.sort(b[f])), (r, f)
        lib.close()
