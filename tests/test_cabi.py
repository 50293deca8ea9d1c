"""CPU-side checks of the drop-in boundary: the CUDA shared library builds, loads and exports every symbol that
include/ruckig_b200.h declares (no compute calls — there is no GPU here), and fails loudly without a device."""
import ctypes
import os
import re
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    sys.path.insert(0, ROOT)
    import __graft_entry__ as g
    g.build()
    return ctypes.CDLL(g.LIB)


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "ruckig_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(rk_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_are_exported(lib):
    syms = declared_symbols()
    assert len(syms) >= 14
    for s in syms:
        assert hasattr(lib, s), f"{s} is declared in include/ruckig_b200.h but not exported"


def test_python_binding_lists_the_same_symbols():
    from rsruckig_b200 import _cabi
    assert sorted(_cabi.EXPORTS) == declared_symbols()


def test_struct_layouts_match_the_header():
    """sizeof of the ctypes mirrors == sizeof of the C structs, measured by compiling a tiny C program."""
    from rsruckig_b200 import _cabi
    src = '#include <stdio.h>\n#include "ruckig_b200.h"\nint main(){printf("%zu %zu %zu %zu %zu\\n", sizeof(rk_batch_desc), sizeof(rk_batch_in), sizeof(rk_batch_out), sizeof(rk_sample_desc), sizeof(rk_sample_out));return 0;}\n'
    exe = os.path.join(ROOT, "oracle", "_build", "sizes")
    os.makedirs(os.path.dirname(exe), exist_ok=True)
    subprocess.run(["gcc", "-x", "c", "-", "-I", os.path.join(ROOT, "include"), "-o", exe], input=src.encode(), check=True)
    sizes = [int(x) for x in subprocess.run([exe], capture_output=True, check=True).stdout.split()]
    mine = [ctypes.sizeof(c) for c in (_cabi.RkBatchDesc, _cabi.RkBatchIn, _cabi.RkBatchOut, _cabi.RkSampleDesc, _cabi.RkSampleOut)]
    assert sizes == mine


def test_no_cpu_fallback(lib):
    """Without a CUDA device rk_create must fail (RK_ERR_NO_DEVICE), not fall back to anything."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    h = ctypes.c_void_p()
    lib.rk_create.argtypes = [ctypes.c_int32, ctypes.POINTER(ctypes.c_void_p)]
    lib.rk_create.restype = ctypes.c_int32
    assert lib.rk_create(0, ctypes.byref(h)) == -3
    lib.rk_last_error.restype = ctypes.c_char_p
    assert b"no CUDA device" in lib.rk_last_error()


def test_product_does_not_reference_the_oracle():
    """Nothing under rsruckig_b200/ may import, include or link oracle/."""
    for dirpath, _, files in os.walk(os.path.join(ROOT, "rsruckig_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in text.replace("the oracle's", "").replace("(the oracle", "") or f in ("workloads.py", "rk_math.cuh", "rk_core.cuh"), f
                assert not re.search(r"^\s*(from|import)\s+oracle", text, flags=re.M), f
                assert '#include "../../oracle' not in text and "rko_" not in text, f


def test_rust_sys_crate_declares_every_function_of_the_header():
    """bindings/rust/ruckig-b200-sys is source only (no Rust toolchain here): at least its list of functions and the field
    constants must follow the header."""
    src = open(os.path.join(ROOT, "bindings", "rust", "ruckig-b200-sys", "src", "lib.rs")).read()
    declared = sorted(set(re.findall(r"pub fn (rk_[a-z0-9_]+)\s*\(", src)))
    assert declared == declared_symbols()
    header = open(os.path.join(ROOT, "include", "ruckig_b200.h")).read()
    fields = re.search(r"enum \{\s*RK_FIELD_STATUS = 0,(.*?)RK_FIELD_COUNT", header, flags=re.S).group(0)
    fields = re.sub(r"/\*.*?\*/", "", fields, flags=re.S)
    names = re.findall(r"\b(RK_FIELD_[A-Z0-9_]+)\b", fields)[:-1]
    for value, name in enumerate(names):
        assert re.search(rf"pub const {name}: i32 = {value};", src), name


def test_library_carries_both_instantiations_of_the_pipeline():
    """rk_kernels.cu compiles the device headers twice (namespace rk: exact divisions on the spot; rk_lazy: flag and redo,
    DESIGN.md 4); the three kernels launched from rk_lazy and their exact counterparts must both be in the library."""
    from rsruckig_b200 import _cabi
    data = open(_cabi.LIB_PATH, "rb").read()
    for sym in (b"_ZN7rk_lazy15k1_closed_candsILi2EEEvNS_6ParamsE", b"_ZN7rk_lazy8k2_stageILi0EEEvNS_6ParamsEii",
                b"_ZN7rk_lazy8k2_stageILi3EEEvNS_6ParamsEii", b"_ZN2rk8k2_stageILi2EEEvNS_6ParamsEii", b"_ZN2rk16k1_quartic_rootsILi0EEEvNS_6ParamsE"):
        assert sym in data, sym.decode()
