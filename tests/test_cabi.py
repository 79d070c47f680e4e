"""The C-ABI library loads and exports every symbol include/rstsr_cuda.h declares; entry points fail loudly
(never fall back to a CPU path) when no CUDA device is present.  Runs without a GPU."""
import ctypes
import os
import re

import pytest

import rstsr_b200 as rt
from rstsr_b200 import _ffi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "rstsr_cuda.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(rc_[a-z0-9_]+)\s*\(", text)))


def test_header_declares_the_expected_surface():
    syms = declared_symbols()
    for must in ("rc_assign", "rc_assign_arbitary", "rc_fill", "rc_op_mutc_refa_refb", "rc_op_muta_refb",
                 "rc_unary_muta_refb", "rc_reduce_all", "rc_reduce_axes", "rc_layout_for_binary_op",
                 "rc_comm_all_reduce", "rc_malloc", "rc_device_create"):
        assert must in syms
    assert len(syms) >= 50


def test_every_declared_symbol_is_exported_and_bound():
    lib = _ffi.lib()
    for name in declared_symbols():
        assert hasattr(lib, name), f"{name} declared in include/rstsr_cuda.h but not exported"
        assert name in _ffi.SIGNATURES, f"{name} has no ctypes signature in rstsr_b200/_ffi.py"
    for name in _ffi.SIGNATURES:
        assert name in declared_symbols(), f"{name} bound in _ffi.py but not declared in the header"


def test_version_and_dtype_sizes():
    lib = _ffi.lib()
    assert b"sm_100a" in lib.rc_version()
    sizes = [lib.rc_dtype_size(t) for t in range(11)]
    assert sizes == [1, 1, 2, 4, 8, 1, 2, 4, 8, 4, 8]


def test_no_cpu_fallback_without_a_device():
    """On a box without CUDA every device entry point must return DeviceError, not compute on the host."""
    lib = _ffi.lib()
    n = ctypes.c_int(0)
    st = lib.rc_device_count(ctypes.byref(n))
    if st == 0 and n.value > 0:
        pytest.skip("a CUDA device is present")
    with pytest.raises(rt.RstsrCudaError) as e:
        rt.DeviceCuda(0)
    assert e.value.kind == "DeviceError"


def test_product_never_imports_the_oracle():
    """`oracle/` is test infrastructure: nothing under rstsr_b200/ may import, load or execute it."""
    pkg = os.path.join(ROOT, "rstsr_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".hpp", ".h")):
                src = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "import oracle" not in src and "from oracle" not in src and "liboracle" not in src, f


def test_rust_ffi_block_is_generated_from_the_header():
    """crates-device/rstsr-cuda/src/ffi.rs must equal what scripts/gen_rust_ffi.py derives from include/rstsr_cuda.h,
    and declare exactly the symbols the ctypes stub binds (header == ctypes stub == Rust extern block)."""
    import importlib.util
    import re
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location("gen_rust_ffi", os.path.join(root, "scripts", "gen_rust_ffi.py"))
    gen = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(gen)
    committed = open(os.path.join(root, "crates-device", "rstsr-cuda", "src", "ffi.rs")).read()
    assert committed == gen.generate(), "run `python scripts/gen_rust_ffi.py` after changing the header"
    from rstsr_b200 import _ffi
    assert set(re.findall(r"pub fn (rc_\w+)\(", committed)) == set(_ffi.SIGNATURES)
