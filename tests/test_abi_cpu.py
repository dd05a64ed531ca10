"""The C-ABI library loads without a GPU and exports every symbol include/zkgpu.h declares; it must refuse to run
(not fall back) when there is no CUDA device."""
import ctypes
import os
import re

import pytest

from era_zkevm_test_harness_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    hdr = open(os.path.join(ROOT, "include", "zkgpu.h")).read()
    return sorted(set(re.findall(r"ZKGPU_API[^;(]*?\b(zkgpu_[a-z0-9_]+)\s*\(", hdr)))


def test_header_and_binding_agree():
    assert _declared_symbols() == sorted(_lib.SYMBOLS)


def test_library_exports_every_declared_symbol():
    if not os.path.exists(_lib.LIB_PATH):
        import __graft_entry__
        __graft_entry__.build()
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for name in _declared_symbols():
        assert hasattr(lib, name), name
    assert _lib.load().zkgpu_abi_version() >= 1


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    lib = _lib.load()
    h = ctypes.c_void_p()
    rc = lib.zkgpu_ctx_create(0, None, ctypes.byref(h))
    assert rc != 0 and b"no CUDA device" in lib.zkgpu_last_error()
    from era_zkevm_test_harness_b200 import GpuContext
    with pytest.raises(_lib.ZkGpuError):
        GpuContext(0)


def test_host_alloc_needs_a_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    lib = _lib.load()
    assert not lib.zkgpu_host_alloc(4096) and b"host_alloc" in lib.zkgpu_last_error()
    lib.zkgpu_host_free(None)
