"""The C-ABI library loads on a CPU-only box and exports every symbol that
include/hssb200.h declares (no compute calls)."""
import os
import re

import pytest


def test_header_symbols_exported(hb):
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    hdr = open(os.path.join(root, "include", "hssb200.h")).read()
    declared = set(re.findall(r"\b(hssb_[a-z0-9_]+)\s*\(", hdr))
    assert len(declared) >= 25
    L = hb.lib()
    for name in sorted(declared):
        assert hasattr(L, name), f"{name} declared in hssb200.h but not exported"
    assert declared == set(hb.SIGNATURES), declared ^ set(hb.SIGNATURES)
    assert L.hssb_version() == 100


def test_errors_are_reported_not_thrown(hb):
    L = hb.lib()
    assert L.hssb_builder_create(None) == -1
    assert b"NULL" in L.hssb_last_error()
    assert L.hssb_device_count() >= 0


def test_library_is_sm100a_with_fp64_tensor_ops(hb):
    """cuobjdump: the shipped .so carries sm_100a SASS with DMMA (FP64 tensor) ops."""
    import shutil
    import subprocess
    if not shutil.which("cuobjdump"):
        return
    out = subprocess.run(["cuobjdump", "-lelf", hb.LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in out
    sass = subprocess.run(["cuobjdump", "-sass", hb.LIB_PATH], capture_output=True, text=True).stdout
    assert "DMMA" in sass


def test_plain_c_client_compiles_and_runs(hb, tmp_path):
    """include/hssb200.h is valid C99 and the library links into a plain C program (the drop-in boundary is a
    C ABI, not a C++ or torch interface); host-only calls, no GPU needed."""
    import shutil
    import subprocess
    if shutil.which("gcc") is None:
        pytest.skip("gcc not available")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = str(tmp_path / "c_abi_smoke")
    libdir = os.path.dirname(hb.LIB_PATH)
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Wextra", "-Werror", "-pedantic", "-I", os.path.join(root, "include"),
                           os.path.join(root, "tests", "c_abi_smoke.c"), "-o", exe, "-L", libdir, "-lhssb200",
                           "-Wl,-rpath," + libdir])
    out = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, (out.returncode, out.stdout, out.stderr)
    assert "C_ABI_OK" in out.stdout
