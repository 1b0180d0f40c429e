import os
import shutil
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def _has_cuda_device():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def _has_nvcc():
    return bool(os.environ.get("NVCC") or shutil.which("nvcc") or os.path.exists("/usr/local/cuda/bin/nvcc"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (runs on the B200 box)")
    config.addinivalue_line("markers", "needs_product_lib: needs libobgpu_scan.so (nvcc) but no device")


def pytest_collection_modifyitems(config, items):
    """`-m gpu` tests are skipped (not failed) on a box without a CUDA device; tests that only need the built
    product library are skipped when neither the library nor nvcc is there. The pure-CPU suites (oracle KATs,
    writer round trips, fuzz twin) need gcc / g++ only."""
    have_dev = _has_cuda_device()
    lib = os.path.join(ROOT, "oceanbase_b200", "csrc", "libobgpu_scan.so")
    have_lib = os.path.exists(lib) or _has_nvcc()
    for it in items:
        if "gpu" in it.keywords and not have_dev:
            it.add_marker(pytest.mark.skip(reason="no CUDA device"))
        elif "needs_product_lib" in it.keywords and not have_lib:
            it.add_marker(pytest.mark.skip(reason="libobgpu_scan.so not built and no nvcc"))


@pytest.fixture(scope="session", autouse=True)
def _built_libraries():
    """CPU side always: the writer library (g++) and the oracle (gcc). The CUDA product library and the C++ adapter
    test binaries are built when nvcc is present (no-ops when up to date); without nvcc a prebuilt library is used
    as is and the tests that need it are skipped when it is absent."""
    import __graft_entry__ as g
    if _has_nvcc():
        g.build()
    else:
        g.build_cpu_side()
