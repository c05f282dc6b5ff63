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
    assert ctypes.sizeof(L.Solver) == 4 * 14 and L.Solver.alg.offset == 52
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


# ------------------------------------------------------------------ host logic of the backward launch plan (no GPU)
def _plan(m, sizes, act, nvars, naug, ncond, B, sm_count=148, exact=0):
    import ctypes as C
    L = m._lib
    cfg = L.Config()
    cfg.abi_version = L.ICNF_ABI_VERSION
    cfg.nvars, cfg.naug, cfg.ncond = nvars, naug, ncond
    cfg.autonomous = int(sizes[0] == nvars + naug + ncond)
    cfg.n_layers = len(sizes) - 1
    for i, s in enumerate(sizes):
        cfg.sizes[i] = s
    cfg.activation = L.ACT[act]
    cfg.precision = L.PRECISION["fp32"]
    threads, grid, nb = C.c_int32(), C.c_int32(), C.c_int32()
    first = (C.c_int32 * 34)()
    rc = m.lib.icnf_backward_plan(C.byref(cfg), exact, sm_count, B, C.byref(threads), C.byref(grid), first, C.byref(nb))
    return rc, threads.value, grid.value, list(first[: nb.value + 1])


@pytest.mark.parametrize("B", [1, 31, 300, 4096, 65536, 200_003, 1_000_000])
@pytest.mark.parametrize("exact", [0, 1])
def test_backward_plan_covers_the_batch(m, B, exact):
    """Config 2's shape (3-12-12-2 softplus): the plan is a pure function of (B, SM count)."""
    rc, threads, grid, first = _plan(m, (3, 12, 12, 2), "softplus", 2, 0, 0, B, exact=exact)
    assert rc == 0
    assert threads % 32 == 0 and 32 <= threads <= 256
    tiles = -(-B // threads)
    assert 1 <= grid <= min(tiles, 2 * 148)                  # at most two CTAs per SM, never more CTAs than tiles
    rounds = -(-tiles // grid)
    assert rounds == -(-B // (2 * 148 * threads)) or threads < 224 or rounds >= 1
    # dW blocks: contiguous thread ranges, every block has threads, nothing beyond the CTA
    assert first[0] == 0 and all(b > a for a, b in zip(first, first[1:])) and first[-1] <= threads
    if threads // 32 >= len(first) - 1:                      # one warp or more per block: a warp never mixes blocks
        assert all(f % 32 == 0 for f in first)


def test_backward_plan_headline_batch_is_one_round(m):
    rc, threads, grid, first = _plan(m, (3, 12, 12, 2), "softplus", 2, 0, 0, 65536)
    assert rc == 0 and threads == 224 and grid == 293         # 293 x 224 >= 65 536: every SM holds its share at once
    assert len(first) - 1 == 7 and first[-1] == 224           # seven dW blocks, one warp each


def test_backward_plan_unknown_shape_is_unsupported(m):
    rc, *_ = _plan(m, (3, 64, 64, 2), "softplus", 2, 0, 0, 1024)
    assert rc == 7                                            # ICNF_ERR_UNSUPPORTED: served by the generic family


def test_steer_tspan_matches_the_draw_spec():
    """icnf_steer_tspan (host function of the library, runs without a GPU) against the oracle's Philox draw
    (steer_tspan, /root/reference/src/core/base_icnf.jl:23-43): only TrainMode{true} steers, rate 0 does not."""
    import ctypes as C
    import numpy as np
    import cnf_b200 as m
    from oracle import icnf_oracle as O
    from oracle import philox as P
    out = C.c_float()
    for seed in (0, 1, 12345, 2 ** 40 + 7, 2 ** 63 - 1):
        for rate, t0, t1 in ((0.1, 0.0, 1.0), (0.25, 0.5, 2.0), (0.1, 1.0, 0.0)):
            assert m.lib.icnf_steer_tspan(1, t0, t1, rate, seed, C.byref(out)) == 0
            om = O.OracleICNF(nvars=1, tspan=(t0, t1), steer_rate=rate)
            ref = O.steer_t1(om, O.TRAIN_REG, P.uniform_pm(seed, rate))
            assert abs(out.value - ref) <= 2e-7 * max(1.0, abs(ref)), (seed, rate, out.value, ref)
            assert abs(out.value - t1) <= rate * abs(t1 - t0) * (1 + 1e-6)
            for mode in (0, 2):
                assert m.lib.icnf_steer_tspan(mode, t0, t1, rate, seed, C.byref(out)) == 0 and out.value == np.float32(t1)
        assert m.lib.icnf_steer_tspan(1, 0.0, 1.0, 0.0, seed, C.byref(out)) == 0 and out.value == 1.0
