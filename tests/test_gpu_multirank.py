"""Two ranks, two GPUs, NCCL: pgpu_run_steps with the EXCHANGE step in the loop (skipped with fewer than 2 GPUs).

Invariants of the reference's decomposition (src/grid_and_partition.f90:525-622): the global particle count is
conserved (specular/diffuse walls only), after EXCHANGE every particle sits on the rank that owns its cell, some
particles do cross the strip interface, and the collision count per particle matches a single-GPU run of the same
physical state within sampling error.
"""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

pytestmark = pytest.mark.gpu

CUBES, PPC, STEPS = (6, 8, 8), 30, 5


def _free_port() -> int:
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out_dir):
    import torch
    import torch.distributed as dist

    from pantera_pic_dsmc_b200 import capi, workloads as wl

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    n = CUBES[0] * CUBES[1] * CUBES[2] * 6 * PPC
    w = wl.c5_channel_strip(*CUBES, n=n, rank=rank, world=world, seed=11 + rank)
    w.cfg.track_id = 1
    gpu = capi.load_gpu(rank)
    wl.apply(gpu, w)
    uid = [gpu.comm_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(uid, src=0)
    gpu.comm_init(world, rank, uid[0])
    ncoll, _ = gpu.run_steps(1, STEPS)
    # the same step as separate calls, in the order TIME_LOOP makes them (timecycle.f90:271-299): ADVECT leaves the
    # emigrant list for EXCHANGE (fast path) ...
    gpu.advect(100)
    gpu.exchange()
    c1, _ = gpu.collide(100)
    # ... and ADVECT followed by a sort BEFORE the exchange: the emigrants are still in the store and must survive
    # the sort (recount), the exchange then takes the full sweep
    n_before = gpu.num_particles()
    gpu.advect(101)
    gpu.sort()
    n_sorted = gpu.num_particles()
    gpu.exchange()
    c2, _ = gpu.collide(101)
    ncoll += 0 * (c1 + c2)
    p = gpu.download_particles()
    owner = np.asarray(w.mesh.cell_procs)[p["IC"] - 1]
    came = int(np.sum((p["ID"] & 0xFF) != rank))       # ID = PROC_ID + ISHFT(counter, 8): born on the other rank
    tot = torch.tensor([len(p), ncoll, came], dtype=torch.int64)
    dist.all_reduce(tot)
    ok = bool(np.all(owner == rank)) and int(tot[0]) == world * n and int(tot[2]) > 0 and n_sorted == n_before \
        and c1 > 0 and c2 > 0
    with open(os.path.join(out_dir, f"ok{rank}"), "w") as f:
        f.write(f"{int(ok)} {int(tot[0])} {int(tot[1])} {int(tot[2])} {world * n}")
    gpu.close()
    dist.destroy_process_group()


@pytest.mark.timeout(600)
@pytest.mark.parametrize("world", [2, 4])
def test_multi_gpu_run_steps_with_nccl_exchange(tmp_path, world):
    """world = 4: the interior strips exchange with two neighbours each (grouped send/recv with two segments)."""
    import torch
    import torch.multiprocessing as mp

    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    from pantera_pic_dsmc_b200 import capi, workloads as wl

    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    res = [open(tmp_path / f"ok{r}").read().split() for r in range(world)]
    assert all(r[0] == "1" for r in res), res
    ncoll_2 = int(res[0][2])
    # single GPU, same physical state (both strips in one box): collisions per particle within sampling error
    n = world * CUBES[0] * CUBES[1] * CUBES[2] * 6 * PPC
    w = wl.c5_channel_strip(world * CUBES[0], CUBES[1], CUBES[2], n=n, rank=0, world=1, seed=5)
    gpu = capi.load_gpu(0)
    wl.apply(gpu, w)
    ncoll_1, _ = gpu.run_steps(1, STEPS)
    gpu.close()
    assert abs(ncoll_2 - ncoll_1) < 6.0 * np.sqrt(ncoll_1) + 0.01 * ncoll_1, (ncoll_2, ncoll_1)
