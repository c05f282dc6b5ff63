"""Parity on the BASELINE.json shapes that bench.py times ("benchmarked but untested"):

  config 4   784-D FFJORD, 785-512-512-512-784 softplus     fp32, split bf16 (1e-4) and plain bf16 (2e-2)
  config 5   CondICNF 64-D + 32 conditions, default width 97-388-388-64, incl. generate
  config 2'  two-moons at the other width SURVEY 8(d) names, 3-64-64-2

Every comparison is the CUDA path through the C ABI against the float64 oracle on the same seeded
inputs, weights and noise.  Tolerances: 1e-4 (north_star) for fp32 and split-bf16 log-densities and
RHS rows; 2e-4 for their gradients; 2e-2 for plain bf16 (8 mantissa bits)."""
import numpy as np
import pytest
import torch

from oracle import icnf_oracle as O
from tests.helpers import make_inputs, norm_rel_err, t64

pytestmark = pytest.mark.gpu

TOL = {"fp32": 1e-4, "bf16x3_tc": 1e-4, "bf16_tc": 2e-2}
GTOL = {"fp32": 2e-4, "bf16x3_tc": 2e-4, "bf16_tc": 3e-2}


@pytest.fixture(scope="module")
def m():
    import cnf_b200
    return cnf_b200


def _ffjord(m, **kw):
    nn = m.Chain(m.Dense(785, 512, "softplus"), m.Dense(512, 512, "softplus"), m.Dense(512, 512, "softplus"), m.Dense(512, 784))
    return m.ICNF(nvariables=784, naugments=0, nn=nn, **kw)


def _cond(m, **kw):
    return m.ICNF(nvariables=64, naugments=0, nconditions=32, **kw)      # default n_hidden = 4 n_in = 388


def _pixels(B, seed):
    return np.random.default_rng(seed).uniform(0.0, 1.0, size=(784, B)).astype(np.float32)   # U[0,1) "pixels"


# ---------------------------------------------------------------- config 4
@pytest.mark.parametrize("prec", ["fp32", "bf16x3_tc", "bf16_tc"])
def test_config4_rhs(m, prec):
    icnf = _ffjord(m, precision=prec)
    assert icnf.sizes == (785, 512, 512, 512, 784)
    assert icnf.kernel_family == ("generic" if prec == "fp32" else "tc")
    B = 300                                            # ragged against every tile size (128 / 64 / 32)
    om, theta, _, eps, _ = make_inputs(icnf, B)
    u = np.concatenate([_pixels(B, 3), np.random.default_rng(4).standard_normal((3, B)).astype(np.float32)])
    for mode, omode in ((m.TrainMode(True), O.TRAIN_REG), (m.TrainMode(False), O.TRAIN_NOREG)):
        du = m.augmented_f(icnf, mode, u, theta, 0.37, eps=eps)
        ref = O.rhs_closed(om, omode, t64(u), t64(theta), 0.37, t64(eps)).numpy()
        for r0, r1 in ((0, om.d), (om.d, om.d + 1), (om.d + 1, om.n_state)):
            err = norm_rel_err(du[r0:r1], ref[r0:r1])
            assert err < TOL[prec], (prec, mode, r0, err)


@pytest.mark.parametrize("prec", ["fp32", "bf16x3_tc", "bf16_tc"])
def test_config4_fixed_step_logp_and_loss(m, prec):
    icnf = _ffjord(m, precision=prec)
    B = 256
    om, theta, _, eps, _ = make_inputs(icnf, B)
    xs = _pixels(B, 5)
    sol = dict(adaptive=False, dt=0.5)
    logp, regs = m.inference(icnf, m.TrainMode(True), xs, theta, {}, eps=eps, tspan=icnf.tspan, **sol)
    ref, rr = O.inference(om, O.TRAIN_REG, t64(xs), t64(theta), t64(eps), opts=O.SolverOpts(adaptive=False, dt=0.5))
    assert norm_rel_err(logp, ref.numpy()) < TOL[prec], norm_rel_err(logp, ref.numpy())
    assert norm_rel_err(regs[0], rr[0].numpy()) < 10 * TOL[prec]
    assert norm_rel_err(regs[1], rr[1].numpy()) < 10 * TOL[prec]


@pytest.mark.parametrize("prec", ["fp32", "bf16x3_tc"])
def test_config4_adaptive_logp(m, prec):
    """adaptive Tsit5 at the reference tolerances: solver-tolerance agreement with the float64 oracle"""
    icnf = _ffjord(m, precision=prec)
    B = 256
    om, theta, _, eps, _ = make_inputs(icnf, B)
    xs = _pixels(B, 6)
    logp, _ = m.inference(icnf, m.TrainMode(False), xs, theta, {}, eps=eps, tspan=icnf.tspan)
    ref, _ = O.inference(om, O.TRAIN_NOREG, t64(xs), t64(theta), t64(eps))
    assert norm_rel_err(logp, ref.numpy()) < 1e-4, norm_rel_err(logp, ref.numpy())
    assert icnf.last_stats.status == 0 and icnf.last_stats.naccept >= 1


@pytest.mark.parametrize("prec", ["fp32", "bf16x3_tc", "bf16_tc"])
def test_config4_loss_gradient(m, prec):
    """RNODE training gradient of the real config-4 shape (1.33 M parameters), theta and xs, against
    reverse-mode AD through the oracle's discrete solve."""
    icnf = _ffjord(m, precision=prec)
    B = 256
    om, theta, _, eps, _ = make_inputs(icnf, B)
    xs = _pixels(B, 7)
    sol = dict(adaptive=False, dt=0.5)
    l, g, gx = m.loss_and_gradient(icnf, m.TrainMode(True), xs, theta, {}, want_dxs=True, eps=eps, tspan=icnf.tspan, **sol)
    rl, rg, rgx = O.loss_grad(om, O.TRAIN_REG, t64(xs), t64(theta), t64(eps), opts=O.SolverOpts(adaptive=False, dt=0.5), want_dxs=True)
    assert abs(l - float(rl)) <= TOL[prec] * abs(float(rl)) + 1e-5
    assert norm_rel_err(g, rg.numpy()) < GTOL[prec], norm_rel_err(g, rg.numpy())
    assert norm_rel_err(gx, rgx.numpy()) < GTOL[prec], norm_rel_err(gx, rgx.numpy())
    # per-layer check: a wrong layer cannot hide behind the largest one
    off = 0
    for nin, nout in zip(icnf.sizes[:-1], icnf.sizes[1:]):
        for n in (nin * nout, nout):
            err = norm_rel_err(g[off:off + n], rg.numpy()[off:off + n])
            assert err < 2 * GTOL[prec], (off, n, err)
            off += n


def test_config4_exact_trace_on_tc_precision_uses_the_fp32_chains(m):
    """TestMode on a 4-layer network has no closed-form trace: D' one-hot chains (utils.jl:35-54) in every precision"""
    icnf = _ffjord(m, precision="bf16x3_tc")
    B = 8
    om, theta, _, _, _ = make_inputs(icnf, B)
    u = np.concatenate([_pixels(B, 8), np.zeros((3, B), np.float32)])
    du = m.augmented_f(icnf, m.TestMode(), u, theta, 0.5)
    ref = O.rhs_closed(om, O.TEST, t64(u), t64(theta), 0.5, None).numpy()
    assert norm_rel_err(du[om.d], ref[om.d]) < 1e-4
    assert norm_rel_err(du[:om.d], ref[:om.d]) < 1e-4


# ---------------------------------------------------------------- config 5 at the default width
@pytest.mark.parametrize("prec", ["fp32", "bf16x3_tc", "bf16_tc"])
def test_config5_default_width(m, prec):
    icnf = _cond(m, precision=prec)
    assert icnf.sizes == (97, 388, 388, 64)
    B = 333
    om, theta, xs, eps, ys = make_inputs(icnf, B)
    theta = (0.5 * theta).astype(np.float32)
    u = np.random.default_rng(3).standard_normal((om.n_state, B)).astype(np.float32)
    for mode, omode in ((m.TrainMode(True), O.TRAIN_REG), (m.TestMode(), O.TEST)):
        du = m.augmented_f(icnf, mode, u, theta, 0.2, eps=eps, ys=ys)
        ref = O.rhs_closed(om, omode, t64(u), t64(theta), 0.2, t64(eps), t64(ys)).numpy()
        for r0, r1 in ((0, om.d), (om.d, om.d + 1)):
            assert norm_rel_err(du[r0:r1], ref[r0:r1]) < TOL[prec], (prec, mode, r0)
    # inference (exact trace, adaptive) and generate (reverse span) through the whole solve
    logp, _ = m.inference(icnf, m.TestMode(), xs, ys, theta, {})
    ref, _ = O.inference(om, O.TEST, t64(xs), t64(theta), None, t64(ys))
    assert norm_rel_err(logp, ref.numpy()) < TOL[prec], norm_rel_err(logp, ref.numpy())
    z0 = np.random.default_rng(5).standard_normal((om.d, B)).astype(np.float32)
    gen = m.generate(icnf, m.TestMode(), ys, theta, {}, B, z0=z0, tspan=icnf.tspan)
    gref = O.generate(om, O.TEST, t64(z0), t64(theta), None, t64(ys)).numpy()
    assert norm_rel_err(gen, gref) < TOL[prec], norm_rel_err(gen, gref)


@pytest.mark.parametrize("prec", ["fp32", "bf16x3_tc"])
def test_config5_loss_gradient(m, prec):
    icnf = _cond(m, precision=prec)
    B = 160
    om, theta, xs, eps, ys = make_inputs(icnf, B)
    theta = (0.5 * theta).astype(np.float32)
    sol = dict(adaptive=False, dt=0.25)
    l, g, gx = m.loss_and_gradient(icnf, m.TrainMode(True), xs, ys, theta, {}, want_dxs=True, eps=eps, tspan=icnf.tspan, **sol)
    rl, rg, rgx = O.loss_grad(om, O.TRAIN_REG, t64(xs), t64(theta), t64(eps), t64(ys), opts=O.SolverOpts(adaptive=False, dt=0.25), want_dxs=True)
    assert abs(l - float(rl)) <= TOL[prec] * abs(float(rl)) + 1e-5
    assert norm_rel_err(g, rg.numpy()) < GTOL[prec], norm_rel_err(g, rg.numpy())
    assert norm_rel_err(gx, rgx.numpy()) < GTOL[prec], norm_rel_err(gx, rgx.numpy())


def test_config5_generate_draws_the_base_sample_in_kernel(m):
    """generate(icnf, mode, ys, ps, st, n) without z0: N(0, I) drawn by Philox inside the solve; the same seed
    reproduces the samples bit for bit and matches the oracle fed with the draw spec's numbers."""
    from oracle import philox as P
    icnf = _cond(m)
    n = 200
    om, theta, _, _, ys = make_inputs(icnf, n)
    theta = (0.5 * theta).astype(np.float32)
    a = m.generate(icnf, m.TestMode(), ys, theta, {}, n, seed=77, tspan=icnf.tspan)
    b = m.generate(icnf, m.TestMode(), ys, theta, {}, n, seed=77, tspan=icnf.tspan)
    assert np.array_equal(a, b)
    z0 = P.gaussian(77, om.d, n, stream=P.STREAM_BASE)
    ref = O.generate(om, O.TEST, t64(z0), t64(theta), None, t64(ys)).numpy()
    assert norm_rel_err(a, ref) < 1e-4, norm_rel_err(a, ref)


# ---------------------------------------------------------------- config 2 at width 64
def test_config2_width64(m):
    icnf = m.ICNF(nvariables=2, naugments=0, n_hidden=64)
    assert icnf.sizes == (3, 64, 64, 2)
    B = 1000
    om, theta, xs, eps, _ = make_inputs(icnf, B)
    for mode, omode in ((m.TrainMode(True), O.TRAIN_REG), (m.TestMode(), O.TEST)):
        logp, regs = m.inference(icnf, mode, xs, theta, {}, eps=eps, tspan=icnf.tspan)
        ref, rr = O.inference(om, omode, t64(xs), t64(theta), t64(eps))
        assert norm_rel_err(logp, ref.numpy()) < 1e-4, (mode, norm_rel_err(logp, ref.numpy()))
    sol = dict(adaptive=False, dt=0.25)
    l, g, gx = m.loss_and_gradient(icnf, m.TrainMode(True), xs, theta, {}, want_dxs=True, eps=eps, tspan=icnf.tspan, **sol)
    rl, rg, rgx = O.loss_grad(om, O.TRAIN_REG, t64(xs), t64(theta), t64(eps), opts=O.SolverOpts(adaptive=False, dt=0.25), want_dxs=True)
    assert abs(l - float(rl)) <= 1e-4 * abs(float(rl)) + 1e-5
    assert norm_rel_err(g, rg.numpy()) < 2e-4, norm_rel_err(g, rg.numpy())
    assert norm_rel_err(gx, rgx.numpy()) < 2e-4


def test_failed_solve_returns_nan_gradient_and_raises(m):
    """a forward solve that cannot finish (max_steps) must not hand a stale or truncated gradient to the
    optimiser -- device path: NaN gradient + check_last() raises; host path: the call raises."""
    icnf = m.ICNF(nvariables=2, naugments=0)
    B = 512
    om, theta, xs, eps, _ = make_inputs(icnf, B)
    with pytest.raises(m.ICNFError):
        m.loss_and_gradient(icnf, m.TrainMode(True), xs, theta, {}, eps=eps, tspan=icnf.tspan, maxiters=1)
    xd, ed, td = (torch.tensor(np.ascontiguousarray(a)).cuda() for a in (xs, eps, theta))
    l, g = m.loss_and_gradient(icnf, m.TrainMode(True), xd, td, {}, eps=ed, tspan=icnf.tspan, maxiters=1)
    assert torch.isnan(g).all()
    with pytest.raises(m.ICNFError):
        icnf.check_last()
    # and a good solve afterwards is clean again
    l, g = m.loss_and_gradient(icnf, m.TrainMode(True), xd, td, {}, eps=ed, tspan=icnf.tspan)
    assert torch.isfinite(g).all() and icnf.check_last().naccept >= 1
