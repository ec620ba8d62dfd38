"""The drop-in boundary without a GPU: header / Python / oracle constants agree, libgsm_b200.so loads and
exports every symbol include/gsm_b200.h declares, and the library fails loudly (no CPU fallback) when no
CUDA device is usable.  No compute entry point is called here."""
import ctypes as C
import os
import re

import pytest

import oracle
from gammaspikemodel_b200 import _lib, abi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = open(os.path.join(ROOT, "include", "gsm_b200.h")).read()
ORACLE_HEADER = open(os.path.join(ROOT, "oracle", "gsm_oracle.h")).read()


def _defines(text, prefix):
    out = {}
    for m in re.finditer(r"#define\s+(%s\w+)\s+\(?\s*(\d+)u?\s*(?:<<\s*(\d+))?\s*\)?" % prefix, text):
        out[m.group(1)] = int(m.group(2)) << int(m.group(3) or 0)
    return out


def test_header_constants_match_python_and_oracle():
    h = _defines(HEADER, "GSM_")
    o = _defines(ORACLE_HEADER, "GSMO_")
    for name, val in h.items():
        if name.startswith(("GSM_F_", "GSM_STUBS_")):
            assert o["GSMO_" + name[4:]] == val, name
            assert getattr(abi, name[4:]) == val, name
            assert getattr(oracle, name[4:]) == val, name
        elif name.startswith(("GSM_P_", "GSM_O_", "GSM_E_")) or name in ("GSM_NPARAM", "GSM_NOUT", "GSM_OK", "GSM_MAX_NODES", "GSM_ABI_VERSION"):
            assert getattr(abi, name[4:]) == val, name
    assert h["GSM_NPARAM"] == o["GSMO_NPARAM"] == 9 and h["GSM_NOUT"] == o["GSMO_NOUT"] == 8


def test_library_exports_every_declared_symbol():
    declared = set(re.findall(r"\b(gsm_[a-z_0-9]+)\s*\(", HEADER))
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    lib = _lib.load()
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.gsm_abi_version() == abi.ABI_VERSION
    assert lib.gsm_device_stride(3999) == 4000 and lib.gsm_device_stride(399) == 416 and lib.gsm_device_stride(1) == 32


def test_no_cpu_fallback_without_a_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    lib = _lib.load()
    assert lib.gsm_device_count() == 0
    h = C.c_void_p()
    rc = lib.gsm_create(C.byref(h), 0, 4, 16)
    assert rc == abi.E_CUDA and not h.value
    assert b"no CPU fallback" in lib.gsm_last_error(None)
    from gammaspikemodel_b200 import GsmEngine, GsmError
    with pytest.raises(GsmError):
        GsmEngine(max_trees=4, max_nodes=16)


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "gammaspikemodel_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".hpp", ".cpp")):
                text = open(os.path.join(dirpath, f)).read()
                assert "gsm_oracle" not in text and "import oracle" not in text and "gsmo_" not in text, f
