"""Particle dump / restart files: DUMP_PARTICLES_FILE / READ_PARTICLES_FILE, binary flavour
(src/tools.f90:327-365, 465-517).

The expected bytes are built here with `struct` from the Fortran I/O statement, independently of the oracle and of
the device: a SEQUENTIAL UNFORMATTED file with CONVERT='BIG_ENDIAN' holds, per WRITE, the record length as a 4-byte
integer, the items, and the length again -- X Y Z VX VY VZ EROT EVIB (real64) S_ID IC (int32) DTRIM (real64) = 80
bytes of payload.  The file name follows the format "(A,A,I0.5,A6,I0.8)".
"""
import os
import struct

import numpy as np
import pytest

from pantera_pic_dsmc_b200 import capi, workloads as wl

import util


def expected_bytes(p):
    out = bytearray()
    for q in p:
        out += struct.pack(">i8d2idi", 80, q["X"], q["Y"], q["Z"], q["VX"], q["VY"], q["VZ"], q["EROT"], q["EVIB"],
                           int(q["S_ID"]), int(q["IC"]), q["DTRIM"], 80)
    return bytes(out)


def workload(n=5000, rank=0):
    w = wl.c2_air_box(n, nx=4, ny=3, with_reactions=False)
    rng = np.random.default_rng(8)
    w.particles["DTRIM"] = rng.random(n) * w.cfg.dt          # freshly injected particles carry their own DTRIM
    w.cfg.rank = rank
    w.cfg.nranks = max(rank + 1, 1)
    w.cfg.track_id = 1
    return w


def roundtrip(lib, tmp_path, rank):
    w = workload(rank=rank)
    wl.apply(lib, w)
    prefix = str(tmp_path) + "/dumps_"
    lib.dump_particles_file(prefix, 1234)
    name = f"{prefix}proc_{rank:05d}_time_{1234:08d}"
    assert os.path.exists(name)
    data = open(name, "rb").read()
    assert len(data) == 88 * len(w.particles)
    assert data == expected_bytes(w.particles), "dump file differs from the reference's record layout"
    with pytest.raises(capi.PanteraError):                   # STATUS='NEW': an existing file is an error
        lib.dump_particles_file(prefix, 1234)
    # restart: append to an empty store; the reference gives every restarted particle DTRIM = DT (tools.f90:497)
    lib.upload_particles(w.particles[:0])
    lib.read_particles_file(prefix, 1234)
    back = lib.download_particles()
    assert len(back) == len(w.particles)
    for f in ("X", "Y", "Z", "VX", "VY", "VZ", "EROT", "EVIB", "IC", "S_ID"):
        assert np.array_equal(back[f], w.particles[f]), f
    assert np.all(back["DTRIM"] == w.cfg.dt)
    # a second read appends (ADD_PARTICLE_ARRAY)
    lib.read_particles_file(prefix, 1234)
    assert lib.num_particles() == 2 * len(w.particles)
    with pytest.raises(capi.PanteraError):
        lib.read_particles_file(prefix, 99)                  # missing restart file


@pytest.mark.parametrize("rank", [0, 3])
def test_oracle_dump_and_restart_follow_the_fortran_record_layout(oracle, tmp_path, rank):
    roundtrip(oracle, tmp_path, rank)


@pytest.mark.gpu
@pytest.mark.parametrize("rank", [0, 3])
def test_device_dump_and_restart_follow_the_fortran_record_layout(gpu, tmp_path, rank):
    roundtrip(gpu, tmp_path, rank)


@pytest.mark.gpu
def test_device_dump_after_a_step_equals_oracle_dump_of_the_same_store(gpu, oracle_philox, tmp_path):
    """After ADVECT with removals the device store holds tombstones; the dump must contain the live particles only, and
    a multi-chunk store (> 2^20 particles) goes through both staging buffers."""
    w = wl.c3_tri_box(n=1_200_000, nx=12, ny=10, wall_bc=("vacuum", "specular", "specular", "specular"), u_drift=(-900.0, 0, 0), dt=2e-5)
    w.cfg.track_id = 1
    wl.apply(gpu, w)
    gpu.advect(1)
    gpu.dump_particles_file(str(tmp_path) + "/g_", 7)
    data = open(f"{tmp_path}/g_proc_00000_time_00000007", "rb").read()
    live = gpu.download_particles()
    assert 0 < len(live) < len(w.particles) and len(data) == 88 * len(live)
    assert data == expected_bytes(live)
    wl.apply(oracle_philox, w, upload=False)
    oracle_philox.upload_particles(live)
    oracle_philox.dump_particles_file(str(tmp_path) + "/o_", 7)
    assert open(f"{tmp_path}/o_proc_00000_time_00000007", "rb").read() == data
