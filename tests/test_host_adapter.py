"""C++ host adapter (oceanbase_b200/host: ObGpuMicroBlockDecoder, ObPushdownFilterExecutor::execute,
ObGpuSSTableBatchScanner) -- tests/cpp/test_host_adapter.cpp is written in the shape of the
reference's decoder tests and runs against the GPU; without a device it must refuse (exit 77)."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "tests", "cpp", "test_host_adapter")


def _has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def test_adapter_builds_and_refuses_without_device():
    assert os.path.exists(BIN)  # built by __graft_entry__.build() (conftest session fixture)
    if _has_gpu():
        pytest.skip("device present: covered by the gpu test")
    r = subprocess.run([BIN], capture_output=True, text=True, timeout=120)
    assert r.returncode == 77, r.stdout + r.stderr


@pytest.mark.gpu
def test_adapter_parity_on_gpu():
    r = subprocess.run([BIN], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout[-4000:] + r.stderr[-2000:]
    assert "host adapter tests passed" in r.stdout


MERGER_BIN = os.path.join(ROOT, "tests", "cpp", "test_partition_merger")


def test_partition_merger_builds_and_refuses_without_device():
    assert os.path.exists(MERGER_BIN)
    if _has_gpu():
        pytest.skip("device present: covered by the gpu test")
    r = subprocess.run([MERGER_BIN], capture_output=True, text=True, timeout=120)
    assert r.returncode == 77, r.stdout + r.stderr


@pytest.mark.gpu
def test_partition_merger_parity_on_gpu():
    r = subprocess.run([MERGER_BIN], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout[-4000:] + r.stderr[-2000:]
    assert "partition merger tests passed" in r.stdout
