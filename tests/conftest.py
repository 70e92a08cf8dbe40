"""pytest wiring: the `gpu` marker and the three library fixtures.

  orc  -- oracle/_build/libmanta_oracle.so, the C restatement (built on demand with `make -C oracle`)
  ref  -- oracle/_ref/libmanta_ref.so, the unmodified reference compiled by oracle/build_ref.sh
          (present in the build container; tests that need it are skipped where it is absent)
  gpu  -- manta_b200.Context on cuda:0 through the C ABI of libmanta_b200.so (tests marked @pytest.mark.gpu)
"""
import ctypes as C
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from manta_b200 import _cabi as cb  # noqa: E402


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on the B200 box)")


@pytest.fixture(scope="session")
def orc():
    lib = os.path.join(ROOT, "oracle", "_build", "libmanta_oracle.so")
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle")])
    return cb.Binding(C.CDLL(lib), "orc_")


@pytest.fixture(scope="session")
def ref():
    lib = os.path.join(ROOT, "oracle", "_ref", "libmanta_ref.so")
    if not os.path.exists(lib):
        if os.path.isdir("/root/reference/src"):
            subprocess.check_call([os.path.join(ROOT, "oracle", "build_ref.sh")])
        else:
            pytest.skip("oracle/_ref/libmanta_ref.so not built (no reference tree on this box)")
    return cb.Binding(C.CDLL(lib), "ref_")


@pytest.fixture(scope="session")
def gpu():
    import manta_b200
    ctx = manta_b200.Context(0)
    yield ctx
    ctx.close()


@pytest.fixture(scope="session")
def gpu32():
    """same library, DP forced onto the general int32 kernel (the default picks the int16x2 pair kernel when the
    score range allows it)"""
    import manta_b200
    ctx = manta_b200.Context(0, force_int32=True)
    yield ctx
    ctx.close()


@pytest.fixture(scope="session")
def gpusel():
    """same library, int16x2 pairs forced onto the predicate/SEL pair kernel (the default is the tagged-score kernel
    where the score range allows it)"""
    import manta_b200
    ctx = manta_b200.Context(0, no_tag_kernel=True)
    yield ctx
    ctx.close()


@pytest.fixture(scope="session", params=["global", "onchip"])
def gpuasm(request):
    """same library with the assembler pinned to one of its two kernels: "global" (window-indexed tables in HBM, the
    throughput kernel of large batches) or "onchip" (dense tables in shared memory, the low-latency kernel the default
    takes for small batches; loci it cannot hold are retried on the other one)"""
    import manta_b200
    ctx = manta_b200.Context(0, assembler=request.param)
    yield ctx
    ctx.close()
