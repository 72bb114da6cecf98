"""The header is a C header: a plain C11 host (what the ISO_C_BINDING interface blocks of INTEGRATION.md bind to, seen from
the C side) must compile against include/pantera_gpu.h with -Wall -Wextra -Werror -pedantic, agree with the ctypes mirror on
the layout of pgpu_config / pgpu_particle / the record planes, link against libpantera_gpu.so referencing EVERY declared
entry point, and -- without a GPU -- get PGPU_ERR_CUDA and a message from pgpu_create (no CPU fallback)."""
import ctypes
import os
import re
import subprocess

import pytest

from pantera_pic_dsmc_b200 import capi

from test_abi import declared_symbols

C_HOST = r"""
#include <stddef.h>
#include <stdio.h>
#include <string.h>
#include "pantera_gpu.h"

_Static_assert(sizeof(pgpu_particle) == 96, "TYPE PARTICLE_DATA_STRUCTURE is 96 bytes (src/particle.f90:28-37)");
_Static_assert(offsetof(pgpu_particle, IC) == 72 && offsetof(pgpu_particle, S_ID) == 76 && offsetof(pgpu_particle, ID) == 80,
               "particle record offsets");
_Static_assert(sizeof(pgpu_config) == 264, "pgpu_config layout");
_Static_assert(sizeof(pgpu_pos_record) == 32 && sizeof(pgpu_vel_record) == 32 && sizeof(pgpu_ie_record) == 16, "record planes");

typedef void (*any_fn)(void);
static const any_fn table[] = {
%(TABLE)s
};

int main(void) {
    pgpu_ctx* ctx = NULL;
    int rc;
    printf("symbols %%d version %%d offsets %%d %%d %%d %%d\n", (int)(sizeof(table) / sizeof(table[0])), (int)pgpu_version(),
           (int)offsetof(pgpu_config, dt), (int)offsetof(pgpu_config, rng_seed_global), (int)offsetof(pgpu_config, capacity),
           (int)offsetof(pgpu_config, track_id));
    rc = pgpu_create(&ctx, 0);
    printf("create %%d\n", rc);
    if (rc == PGPU_OK) {
        int64_t n = -1;
        pgpu_config cfg;
        memset(&cfg, 0, sizeof(cfg));
        cfg.dims = 2; cfg.nx = 1; cfg.ny = 1; cfg.xmax = 1.0; cfg.ymax = 1.0; cfg.zmin = -0.5; cfg.zmax = 0.5; cfg.dt = 1e-6; cfg.fnum = 1.0;
        cfg.nranks = 1;
        rc = pgpu_set_config(ctx, &cfg);
        printf("set_config %%d\n", rc);
        rc = pgpu_num_particles(ctx, &n);
        printf("num_particles %%d %%lld\n", rc, (long long)n);
        pgpu_destroy(ctx);
    }
    return 0;
}
"""


def build_host(tmp_path):
    syms = declared_symbols()
    src = tmp_path / "abi_host.c"
    src.write_text(C_HOST % dict(TABLE=",\n".join(f"    (any_fn){s}" for s in syms)))
    exe = tmp_path / "abi_host"
    lib_dir = os.path.dirname(capi.GPU_LIB_PATH)
    cmd = ["gcc", "-std=c11", "-Wall", "-Wextra", "-Werror", "-pedantic", "-Wno-cast-function-type", "-I", os.path.join(capi.REPO_ROOT, "include"),
           str(src), "-o", str(exe), "-L", lib_dir, "-lpantera_gpu", f"-Wl,-rpath,{lib_dir}"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return exe, len(syms)


def run_host(tmp_path):
    exe, nsym = build_host(tmp_path)
    r = subprocess.run([str(exe)], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, (r.stdout, r.stderr)
    lines = r.stdout.strip().splitlines()
    head = [int(x) for x in re.findall(r"-?\d+", lines[0])]
    assert head[0] == nsym and head[1] >= 1000
    cfg = capi.Config
    assert head[2:] == [getattr(cfg, f).offset for f in ("dt", "rng_seed_global", "capacity", "track_id")]
    return lines


def test_c_host_compiles_links_every_entry_point_and_sees_no_fallback(tmp_path):
    import torch
    lines = run_host(tmp_path)
    if not torch.cuda.is_available():
        assert lines[1] == f"create {2}"            # PGPU_ERR_CUDA: no device, no CPU path behind the ABI


@pytest.mark.gpu
def test_c_host_creates_a_context_on_the_device(tmp_path):
    lines = run_host(tmp_path)
    assert lines[1] == "create 0" and lines[2] == "set_config 0" and lines[3] == "num_particles 0 0"
