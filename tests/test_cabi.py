"""The C-ABI library builds for sm_100a, loads without a GPU and exports every symbol that
include/*.h declares (no compute calls here)."""
import ctypes
import glob
import os
import re

from conftest import ROOT


def declared_symbols():
    names = set()
    for hdr in glob.glob(os.path.join(ROOT, "include", "*.h")):
        text = open(hdr).read()
        text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
        names |= set(re.findall(r"\b(gad_[a-z0-9_]+)\s*\(", text))
    return names


def test_header_symbols_exported(native_lib):
    names = declared_symbols()
    assert len(names) >= 10
    handle = ctypes.CDLL(native_lib.LIB_PATH)
    missing = [n for n in sorted(names) if not hasattr(handle, n)]
    assert not missing, missing


def test_binding_table_matches_header(native_lib):
    assert set(native_lib.SIGNATURES) == declared_symbols()
    lib = native_lib.lib()
    assert lib.gad_version() >= 100


def test_library_has_sm100a_code(native_lib):
    import subprocess
    out = subprocess.run(["cuobjdump", "-lelf", native_lib.LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in out


def test_argument_errors_without_gpu(native_lib):
    """Validation happens before any CUDA call, so it is testable on the CPU box."""
    lib = native_lib.lib()
    rc = lib.gad_fitness(None, None, 4, 0, 16, 0, 4, -4, -4, None, None, None, None, None)
    assert rc == native_lib.GAD_EINVAL
    assert b"R" in lib.gad_last_error()
    rc = lib.gad_fitness(None, None, 4, 3, 10, 0, 4, -4, -4, None, None, None, None, None)
    assert rc == native_lib.GAD_EINVAL and b"stride" in lib.gad_last_error()
