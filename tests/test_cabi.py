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


def header_prototypes():
    """{name: [parameter declarations]} of every gad_* function declared in include/*.h."""
    protos = {}
    for hdr in glob.glob(os.path.join(ROOT, "include", "*.h")):
        text = re.sub(r"/\*.*?\*/", "", open(hdr).read(), flags=re.S)
        for m in re.finditer(r"\b(gad_[a-z0-9_]+)\s*\(([^;{]*?)\)\s*;", text, flags=re.S):
            params = [p.strip() for p in m.group(2).replace("\n", " ").split(",")]
            protos[m.group(1)] = [] if params in (["void"], [""]) else params
    return protos


def test_binding_arity_and_kinds_match_the_header(native_lib):
    """Every ctypes signature has as many arguments as the C prototype, pointers where the header has pointers,
    64-bit integers where it has int64_t / size_t, and floats where it has float - an ABI drift between
    include/gadpamsa.h and _native.py would otherwise only show on the GPU box."""
    protos = header_prototypes()
    assert set(protos) == set(native_lib.SIGNATURES)
    pointer_types = (ctypes.c_void_p, ctypes.c_char_p)
    for name, params in sorted(protos.items()):
        restype, argtypes = native_lib.SIGNATURES[name]
        assert len(argtypes) == len(params), (name, len(argtypes), params)
        for decl, ct in zip(params, argtypes):
            is_ptr = "*" in decl
            ct_is_ptr = ct in pointer_types or hasattr(ct, "contents") or getattr(ct, "_type_", None) == "P"
            assert is_ptr == ct_is_ptr, (name, decl, ct)
            if not is_ptr:
                if re.search(r"\b(int64_t|size_t|long long)\b", decl):
                    assert ctypes.sizeof(ct) == 8, (name, decl, ct)
                elif re.search(r"\bfloat\b", decl):
                    assert ct is ctypes.c_float, (name, decl, ct)
                elif re.search(r"\bdouble\b", decl):
                    assert ct is ctypes.c_double, (name, decl, ct)
                else:
                    assert ctypes.sizeof(ct) == 4, (name, decl, ct)


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


def test_product_path_fails_loudly_without_a_gpu(native_lib):
    """No CPU fallback: on a machine without a CUDA device every compute entry of the facade raises
    (nothing under ga-dpamsa_b200/ may import the oracle or compute on the host)."""
    import pytest
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    import config      # noqa: F401
    import utils
    import GA
    board = [[1, 2, 3], [1, 2, 4]]
    with pytest.raises(native_lib.GadError):
        utils.get_sum_of_pairs(board, 0, 2, 0, 3)
    with pytest.raises(native_lib.GadError):
        utils.clean_unnecessary_gaps([[1, 5], [2, 5]])
    with pytest.raises(native_lib.GadError):
        GA.GA(["ACGT", "ACGA"], "sp").generate_population()
    src = "".join(open(os.path.join(ROOT, "ga-dpamsa_b200", f)).read()
                  for f in os.listdir(os.path.join(ROOT, "ga-dpamsa_b200")) if f.endswith(".py"))
    assert "ga_oracle" not in src and "import oracle" not in src
