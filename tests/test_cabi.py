"""The C-ABI boundary without a GPU: libmeshode_b200.so loads, exports every symbol that
include/meshode_b200.h declares, the ctypes table covers the header, and -- on a box without a
CUDA device -- compute entries fail loudly (no CPU fallback)."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "meshode_b200.h")


def _declared():
    txt = open(HEADER).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(mo_[a-z0-9_]+)\s*\(", txt)))


@pytest.fixture(scope="module")
def lib():
    from meshode_b200 import build, capi
    build.build_lib()
    return capi.lib()


def test_header_is_plain_c(tmp_path):
    src = tmp_path / "t.c"
    src.write_text('#include "meshode_b200.h"\nint main(void){return MO_OK + MO_EDGES_RIGID;}\n')
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), "-c", str(src), "-o",
                           str(tmp_path / "t.o")])
    txt = open(HEADER).read()
    assert "at::Tensor" not in txt and "#include <torch" not in txt


def test_exports_every_declared_symbol(lib):
    names = _declared()
    assert len(names) >= 25
    for n in names:
        assert hasattr(lib, n), "libmeshode_b200.so does not export %s" % n
    out = subprocess.run(["nm", "-D", "--defined-only", lib._name], capture_output=True, text=True).stdout
    exported = set(re.findall(r" T (mo_[a-z0-9_]+)", out))
    assert set(names) <= exported
    assert exported <= set(names), "exported but undeclared: %s" % sorted(exported - set(names))


def test_ctypes_table_covers_header():
    from meshode_b200 import capi
    assert sorted(capi.SIGNATURES) == _declared()


def test_sm100a_only(lib):
    out = subprocess.run(["cuobjdump", "--list-elf", lib._name], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_(\d+a?)", out))
    assert archs == {"100a"}, archs


def test_no_cpu_fallback_without_device(lib):
    from meshode_b200 import capi
    if lib.mo_device_count() > 0:
        pytest.skip("a CUDA device is visible")
    V = np.zeros((3, 3), np.float32); F = np.zeros((1, 3), np.int32)
    pid = C.c_int(-1)
    rc = lib.mo_template_create(V.ctypes.data, 3, F.ctypes.data, 1, 0, 8, None, C.byref(pid))
    assert rc == -3 and pid.value == -1                       # MO_ERR_CUDA
    assert b"CUDA error" in lib.mo_last_error()
    with pytest.raises(capi.MeshodeError):
        capi.require_device()
    with pytest.raises(capi.MeshodeError):
        capi.template_create(V.ctypes.data, 3, F.ctypes.data, 1, 0, 8)
    assert lib.mo_distance_forward(V.ctypes.data, 3, 0, V.ctypes.data, None) == -1   # MO_ERR_BAD_HANDLE
    assert lib.mo_template_destroy(5) == -1


def test_argument_checks(lib):
    pid = C.c_int(-1)
    V = np.zeros((3, 3), np.float32); F = np.zeros((1, 3), np.int32)
    assert lib.mo_template_create(None, 3, F.ctypes.data, 1, 0, 8, None, C.byref(pid)) == -2
    assert lib.mo_template_create(V.ctypes.data, 0, F.ctypes.data, 1, 0, 8, None, C.byref(pid)) == -2
    assert lib.mo_template_create(V.ctypes.data, 3, F.ctypes.data, 1, 0, 1, None, C.byref(pid)) == -2
    assert lib.mo_template_create_slab(V.ctypes.data, 3, F.ctypes.data, 1, 8, 4, 4, None, C.byref(pid)) == -2
    assert b"bad argument" in lib.mo_last_error()
    assert lib.mo_version() >= 1


def test_product_never_imports_oracle():
    """The product package must not reference oracle/ (the oracle is test infrastructure)."""
    pkg = os.path.join(ROOT, "meshode_b200")
    for dp, _, fs in os.walk(pkg):
        for f in fs:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cc")):
                txt = open(os.path.join(dp, f), errors="replace").read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", txt, flags=re.M), os.path.join(dp, f)
                assert "libmeshode_oracle" not in txt and "orc_" not in txt, os.path.join(dp, f)
