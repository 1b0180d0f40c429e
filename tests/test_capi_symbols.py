"""The C-ABI library loads on a CPU-only box and exports every symbol include/*.h
declare (no compute calls without a GPU); the product fails loudly without a device."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.needs_product_lib


WRITER_HEADERS = ("obgpu_writer.h",)


def declared_in_header(writer=False):
    names = set()
    for h in sorted(os.listdir(os.path.join(ROOT, "include"))):
        if not h.endswith(".h") or (h in WRITER_HEADERS) != writer:
            continue
        text = open(os.path.join(ROOT, "include", h)).read()
        text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
        names |= set(re.findall(r"\b(obgpu_[a-z0-9_]+)\s*\(", text))
    return sorted(names)


def test_library_exports_every_declared_symbol():
    import oceanbase_b200 as ob
    L = ctypes.CDLL(ob.lib_path)
    names = declared_in_header()
    assert len(names) >= 30
    for n in names:
        assert hasattr(L, n), f"{n} declared in include/*.h but not exported"
    W = ctypes.CDLL(ob.capi.writer_lib_path)
    wnames = declared_in_header(writer=True)
    assert len(wnames) >= 9
    for n in wnames:
        assert hasattr(W, n), f"{n} declared in include/obgpu_writer.h but not exported by libobgpu_writer.so"
        assert not hasattr(L, n), f"{n}: the writer must not live in the CUDA product library"


def test_python_binding_covers_the_header():
    from oceanbase_b200 import capi
    assert sorted(capi.declared_signatures()) == declared_in_header()
    assert sorted(capi.writer_signatures()) == declared_in_header(writer=True)


def test_writer_library_has_no_cuda_dependency():
    import oceanbase_b200 as ob
    out = os.popen(f"ldd {ob.capi.writer_lib_path}").read()
    assert "cuda" not in out.lower() and "obgpu_scan" not in out


def test_no_cpu_fallback_without_device():
    import oceanbase_b200 as ob
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        pytest.skip("a device is present")
    with pytest.raises(ob.ObGpuError) as ei:
        ob.ScanContext(0)
    assert ei.value.code == ob.OB_ERR_SYS


def test_product_does_not_link_or_import_the_oracle():
    import oceanbase_b200
    pkg = os.path.dirname(oceanbase_b200.__file__)
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                src = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "ob_oracle" not in src and "oracle_binding" not in src, f
    out = os.popen(f"ldd {oceanbase_b200.lib_path}").read()
    assert "oracle" not in out
