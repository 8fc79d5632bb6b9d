"""The C-ABI library loads on a CPU-only box and exports every symbol that
include/hssb200.h declares (no compute calls)."""
import os
import re


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
