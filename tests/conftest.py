"""pytest configuration: the `gpu` marker and shared fixtures.

`-m "not gpu"` tests run on the build container (no GPU): oracle vs known answers, host logic, ABI
surface.  `-m gpu` tests are the parity tests proper: CUDA library vs oracle through the C ABI.
"""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session", autouse=True)
def _build_oracle():
    """The oracle is test infrastructure; build it on demand (gcc only, <2 s)."""
    so = os.path.join(ROOT, "oracle", "_build", "liboracle.so")
    src = os.path.join(ROOT, "oracle", "pantera_oracle.c")
    if not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle")], stdout=subprocess.DEVNULL)
    yield


@pytest.fixture
def oracle():
    from pantera_pic_dsmc_b200 import capi
    from oracle.binding import load_oracle
    lib = load_oracle(rng_mode=0)
    yield lib
    lib.close()


@pytest.fixture
def oracle_philox():
    from pantera_pic_dsmc_b200 import capi
    from oracle.binding import load_oracle
    lib = load_oracle(rng_mode=1)
    yield lib
    lib.close()


@pytest.fixture
def gpu():
    """The CUDA library.  Fails loudly (never falls back) when it is missing or no GPU is visible."""
    from pantera_pic_dsmc_b200 import capi
    lib = capi.load_gpu(0)
    yield lib
    lib.close()
