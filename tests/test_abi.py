"""The C-ABI library must load and export every symbol include/pantera_gpu.h declares (no GPU needed)."""
import ctypes
import os
import re

import pytest

from pantera_pic_dsmc_b200 import capi
from oracle.binding import ORACLE_LIB_PATH

HEADER = os.path.join(capi.REPO_ROOT, "include", "pantera_gpu.h")


def declared_symbols():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(pgpu_[a-z0-9_]+)\s*\(", src)))


def test_header_declares_the_hot_path():
    syms = declared_symbols()
    for s in ["pgpu_create", "pgpu_advect", "pgpu_sort", "pgpu_collide", "pgpu_deposit_charge", "pgpu_exchange",
              "pgpu_mcc_vahedi", "pgpu_run_steps", "pgpu_upload_particles", "pgpu_download_particles"]:
        assert s in syms


def test_library_exports_every_declared_symbol():
    assert os.path.exists(capi.GPU_LIB_PATH), "libpantera_gpu.so missing: run __graft_entry__.build()"
    lib = ctypes.CDLL(capi.GPU_LIB_PATH)
    missing = [s for s in declared_symbols() if not hasattr(lib, s)]
    assert not missing, f"symbols declared in pantera_gpu.h but not exported: {missing}"
    f = lib.pgpu_version
    f.restype = ctypes.c_int
    assert f() >= 1000


def test_oracle_mirrors_the_abi():
    """Every compute/set entry point has a po_ twin so that one driver runs both sides of a parity test."""
    lib = ctypes.CDLL(ORACLE_LIB_PATH)
    skip = {"pgpu_version", "pgpu_comm_unique_id", "pgpu_comm_init", "pgpu_exchange", "pgpu_get_timings"}
    missing = [s for s in declared_symbols() if s not in skip and not hasattr(lib, "po_" + s[5:])]
    assert not missing, missing


def test_config_struct_layout_matches_header():
    """ctypes mirror of pgpu_config must have the C layout (sizes are what a Fortran BIND(C) type gets)."""
    cfg = capi.make_config()
    assert ctypes.sizeof(cfg) == 264, ctypes.sizeof(cfg)  # == sizeof(pgpu_config) (gcc x86-64)
    assert capi.PARTICLE_DTYPE.itemsize == 96


def test_no_cuda_device_is_a_loud_error():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is visible")
    with pytest.raises(capi.PanteraError):
        capi.load_gpu(0)


def test_product_package_never_touches_the_oracle():
    """Mechanical, not a promise: no file of the product package names the oracle library or its binding, importing the
    package pulls in no module of oracle/, and libpantera_gpu.so neither links nor dlopens it."""
    import subprocess
    import sys
    pkg = os.path.dirname(capi.__file__)
    offenders = []
    for root, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", "Makefile")):
                text = open(os.path.join(root, f), errors="replace").read()
                if re.search(r"liboracle|load_oracle|OracleLib|oracle\.binding|from oracle|import oracle|po_create", text):
                    offenders.append(os.path.join(root, f))
    assert not offenders, offenders
    code = "import sys; import pantera_pic_dsmc_b200, pantera_pic_dsmc_b200.capi, pantera_pic_dsmc_b200.workloads; " \
           "print([m for m in sys.modules if m == 'oracle' or m.startswith('oracle.')])"
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, cwd=capi.REPO_ROOT)
    assert out.returncode == 0 and out.stdout.strip() == "[]", (out.stdout, out.stderr)
    needed = subprocess.run(["readelf", "-d", capi.GPU_LIB_PATH], capture_output=True, text=True).stdout
    assert "oracle" not in needed
    strings = subprocess.run(["strings", "-n", "6", capi.GPU_LIB_PATH], capture_output=True, text=True).stdout
    assert "liboracle" not in strings
