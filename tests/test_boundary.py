"""CPU checks of the drop-in boundary: the C-ABI library loads, exports every
symbol include/icnf_b200.h declares, and refuses to run without a GPU."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "icnf_b200.h")


@pytest.fixture(scope="module")
def m():
    lib_path = os.path.join(ROOT, "continuousnormalizingflows.jl_b200", "libicnf_b200.so")
    if not os.path.exists(lib_path):
        import __graft_entry__
        __graft_entry__.build()
    import cnf_b200
    return cnf_b200


def declared_symbols():
    text = open(HEADER).read()
    return sorted(set(re.findall(r"ICNF_API[^;(]*?\b(icnf_\w+)\s*\(", text)))


def test_header_declares_the_expected_entry_points():
    names = declared_symbols()
    for want in ("icnf_create", "icnf_destroy", "icnf_set_params", "icnf_rhs", "icnf_solve", "icnf_inference",
                 "icnf_generate", "icnf_loss", "icnf_loss_grad", "icnf_last_error"):
        assert want in names


def test_library_exports_every_declared_symbol(m):
    cdll = ctypes.CDLL(m.LIB_PATH)
    for name in declared_symbols():
        assert hasattr(cdll, name), f"{name} declared in icnf_b200.h but not exported"
    bound = {s[0] for s in __import__("cnf_b200")._lib.SYMBOLS}
    assert bound == set(declared_symbols())


def test_struct_layouts_match_the_header(m):
    L = m._lib
    assert ctypes.sizeof(L.Config) == 4 * (6 + 9 + 1 + 3 + 3)
    assert ctypes.sizeof(L.Solver) == 4 * 13
    assert ctypes.sizeof(L.Noise) == 24 and L.Noise.seed.offset == 8
    assert ctypes.sizeof(L.Stats) == 24


def test_no_cpu_fallback(m):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    assert m.lib.icnf_device_count() == 0
    with pytest.raises(m.ICNFError) as ei:
        m.ICNF(nvariables=1)
    assert ei.value.code == 6          # ICNF_ERR_NO_DEVICE


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "continuousnormalizingflows.jl_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".jl")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle", src, re.M), f
