"""The C-ABI library builds, loads without a GPU and exports exactly what include/mm3d.h declares."""
import ctypes
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "mm3d.h")


def _declared():
    src = open(HEADER).read()
    return sorted(set(re.findall(r"MM3D_API\s+[\w\s\*]+?\b(mm3d_\w+)\s*\(", src)))


def test_library_builds_and_exports_header_symbols():
    from mm2d3d_b200 import build
    lib_path = build.build()
    assert os.path.exists(lib_path)
    names = _declared()
    assert len(names) >= 20
    lib = ctypes.CDLL(lib_path)
    for n in names:
        assert hasattr(lib, n), f"{n} declared in mm3d.h but not exported"
    exported = subprocess.run(["nm", "-D", "--defined-only", lib_path], capture_output=True, text=True).stdout
    exported = sorted(set(re.findall(r" T (mm3d_\w+)", exported)))
    assert exported == names, "exported symbols and header declarations differ"


def test_python_binding_covers_header():
    from mm2d3d_b200 import _lib
    assert sorted(_lib.SIGNATURES) == _declared()
    assert _lib.lib.mm3d_abi_version() == _lib.ABI_VERSION
    # pure host helpers are callable without a GPU
    assert _lib.lib.mm3d_hash_capacity(1000) >= 2000
    assert _lib.lib.mm3d_unique_workspace_bytes(1000) > 4000
    assert _lib.lib.mm3d_bnrelu_workspace_bytes(16) >= 256


def test_sm100a_code_is_embedded():
    from mm2d3d_b200 import build
    out = subprocess.run(["cuobjdump", "-lelf", build.LIB], capture_output=True, text=True).stdout
    assert "sm_100a" in out
