"""GPU parity tests of the generic kernel family (arbitrary MLP widths, fp32
tiled SGEMMs with fused epilogues, device-resident Tsit5 controller) against the
CPU oracle.  Same tolerance as the tiny family: RTOL = 1e-4 (north_star)."""
import numpy as np
import pytest

from oracle import icnf_oracle as O
from oracle import philox as P
from tests.helpers import GENERIC_SHAPES, make_icnf, make_inputs, t64

pytestmark = pytest.mark.gpu
RTOL = 1e-4


@pytest.fixture(scope="module")
def m():
    import cnf_b200
    return cnf_b200


def modes(m):
    return [(m.TestMode(), O.TEST), (m.TrainMode(True), O.TRAIN_REG), (m.TrainMode(False), O.TRAIN_NOREG)]


@pytest.mark.parametrize("shape", list(GENERIC_SHAPES))
def test_rhs_matches_oracle(m, shape):
    icnf = make_icnf(m, shape)
    assert icnf.kernel_family == "generic"
    B = 131
    om, theta, xs, eps, ys = make_inputs(icnf, B)
    u = np.random.default_rng(3).standard_normal((om.n_state, B)).astype(np.float32)
    for mode, omode in modes(m):
        du = m.augmented_f(icnf, mode, u, theta, 0.37, eps=eps, ys=ys)
        ref = O.rhs_closed(om, omode, t64(u), t64(theta), 0.37, t64(eps), t64(ys)).numpy()
        np.testing.assert_allclose(du, ref, rtol=RTOL, atol=2e-5)


@pytest.mark.parametrize("shape", list(GENERIC_SHAPES))
def test_fixed_step_solve_agrees_step_for_step(m, shape):
    icnf = make_icnf(m, shape)
    om, theta, xs, eps, ys = make_inputs(icnf, 70)
    u0 = O.make_u0(om, t64(xs))
    dt = 0.125
    for mode, omode in modes(m):
        for k in (1, 3, 8):
            got = m.base_sol(icnf, mode, u0.numpy().astype(np.float32), theta, tspan=(0.0, k * dt), eps=eps, ys=ys,
                             adaptive=False, dt=dt)
            ref = O.solve(om, omode, u0, t64(theta), t64(eps), t64(ys), 0.0, k * dt,
                          O.SolverOpts(adaptive=False, dt=dt)).numpy()
            assert icnf.last_stats.naccept == k
            np.testing.assert_allclose(got, ref, rtol=RTOL, atol=3e-5)


@pytest.mark.parametrize("shape", list(GENERIC_SHAPES))
@pytest.mark.parametrize("scale", [1.0, 2.0])
def test_adaptive_inference_matches_oracle(m, shape, scale):
    """See tests/test_gpu_parity.py: adaptive parity is tolerance-level; step counts are only
    compared for the stiffer flow (scale 2), where rounding does not pick the steps."""
    icnf = make_icnf(m, shape)
    om, theta, xs, eps, ys = make_inputs(icnf, 300)
    theta = (scale * theta).astype(np.float32)
    for mode, omode in modes(m):
        args = (xs,) if ys is None else (xs, ys)
        logp, (E, n, A) = m.inference(icnf, mode, *args, theta, {}, eps=eps, tspan=icnf.tspan)
        gs = icnf.last_stats
        st = O.SolveStats()
        rl, (rE, rn, rA) = O.inference(om, omode, t64(xs), t64(theta), t64(eps), t64(ys), stats=st)
        assert gs.status == 0 and gs.t_final == pytest.approx(1.0)
        assert gs.nf == 2 + 6 * (gs.naccept + gs.nreject)
        if scale > 1.0:
            assert abs(gs.naccept - st.naccept) <= 1 and abs(gs.nreject - st.nreject) <= 1, (gs, st.naccept, st.nreject)
        tol = RTOL if scale == 1.0 else 5e-4
        np.testing.assert_allclose(logp, rl.numpy(), rtol=tol, atol=2e-5)
        np.testing.assert_allclose(E, rE.numpy(), rtol=tol, atol=2e-5)
        np.testing.assert_allclose(A, rA.numpy(), rtol=tol, atol=2e-5)
        np.testing.assert_allclose(n, rn.numpy(), rtol=tol, atol=2e-2 * float(rn.abs().mean()) + 1e-5)


def test_generate_and_in_kernel_noise(m):
    icnf = make_icnf(m, "cond_generic", epsdist="rademacher")
    om, theta, xs, eps, ys = make_inputs(icnf, 90)
    z0 = np.random.default_rng(9).standard_normal((om.d, 90)).astype(np.float32)
    got = m.generate(icnf, m.TestMode(), ys, theta, {}, 90, z0=z0, tspan=icnf.tspan)
    ref = O.generate(om, O.TEST, t64(z0), t64(theta), None, t64(ys)).numpy()
    np.testing.assert_allclose(got, ref, rtol=RTOL, atol=3e-5)
    got = m.generate(icnf, m.TestMode(), ys, theta, {}, 90, seed=77, tspan=icnf.tspan)
    zd = P.gaussian(77, om.d, 90, stream=P.STREAM_BASE)
    ref = O.generate(om, O.TEST, t64(zd), t64(theta), None, t64(ys)).numpy()
    np.testing.assert_allclose(got, ref, rtol=RTOL, atol=5e-5)
    logp, (E, n, A) = m.inference(icnf, m.TrainMode(True), xs, ys, theta, {}, seed=5, sample_offset=64,
                                  tspan=icnf.tspan, adaptive=False, dt=0.25)
    er = P.rademacher(5, om.d, 90, offset=64)
    rl, (rE, rn, rA) = O.inference(om, O.TRAIN_REG, t64(xs), t64(theta), t64(er), t64(ys),
                                   opts=O.SolverOpts(adaptive=False, dt=0.25))
    np.testing.assert_allclose(logp, rl.numpy(), rtol=RTOL, atol=2e-5)
    np.testing.assert_allclose(n, rn.numpy(), rtol=RTOL, atol=2e-5)


def test_config3_exact_trace_at_a_larger_batch(m):
    # 16-D GMM data, TestMode: the bilinear closed-form trace against D' autograd pullbacks
    icnf = make_icnf(m, "config3_gmm16")
    om, theta, xs, eps, _ = make_inputs(icnf, 4099)
    logp, _ = m.inference(icnf, m.TestMode(), xs, theta, {}, adaptive=False, dt=0.25, tspan=icnf.tspan)
    sub = slice(0, 64)
    ref, _ = O.inference(om, O.TEST, t64(xs[:, sub]), t64(theta), None, opts=O.SolverOpts(adaptive=False, dt=0.25),
                         closed=False)
    np.testing.assert_allclose(logp[sub], ref.detach().numpy(), rtol=RTOL, atol=2e-5)
    assert np.isfinite(logp).all()


@pytest.mark.parametrize("shape", ["config3_gmm16", "cond_generic"])
@pytest.mark.parametrize("B", [2048, 2051])
def test_weights_stationary_gemm_paths(m, shape, B):
    """Batches >= 2048 route the narrow layers through the weights-stationary persistent SGEMM: vector
    (B % 4 == 0) and scalar tile loads / epilogues, the gathered [z; t; ys] operand, and every epilogue
    the reverse sweep uses."""
    from tests.helpers import norm_rel_err
    icnf = make_icnf(m, shape)
    om, theta, xs, eps, ys = make_inputs(icnf, B)
    args = (xs,) if ys is None else (xs, ys)
    sol = dict(adaptive=False, dt=0.25)
    opts = O.SolverOpts(adaptive=False, dt=0.25)
    for mode, omode in [(m.TrainMode(True), O.TRAIN_REG), (m.TestMode(), O.TEST)]:
        logp, regs = m.inference(icnf, mode, *args, theta, {}, eps=eps, tspan=icnf.tspan, **sol)
        for sub in (slice(0, 40), slice(B - 40, B)):
            ref, rr = O.inference(om, omode, t64(xs[:, sub]), t64(theta), t64(eps[:, sub]) if eps is not None else None,
                                  t64(ys[:, sub]) if ys is not None else None, opts=opts)
            np.testing.assert_allclose(logp[sub], ref.detach().numpy(), rtol=RTOL, atol=2e-5)
    l, g = m.loss_and_gradient(icnf, m.TrainMode(True), *args, theta, {}, eps=eps, tspan=icnf.tspan, **sol)
    rl, rg = O.loss_grad(om, O.TRAIN_REG, t64(xs), t64(theta), t64(eps), t64(ys), opts=opts)[:2]
    assert abs(l - float(rl)) <= RTOL * abs(float(rl)) + 1e-6
    assert norm_rel_err(g, rg.numpy()) < RTOL


@pytest.mark.parametrize("shape", list(GENERIC_SHAPES))
@pytest.mark.parametrize("adaptive", [False, True])
def test_loss_and_gradient_match_oracle(m, shape, adaptive):
    from tests.helpers import norm_rel_err
    icnf = make_icnf(m, shape)
    om, theta, xs, eps, ys = make_inputs(icnf, 77)
    sol = dict(adaptive=False, dt=0.25) if not adaptive else dict(adaptive=True)
    opts = O.SolverOpts(adaptive=False, dt=0.25) if not adaptive else O.SolverOpts()
    for mode, omode in [(m.TrainMode(True), O.TRAIN_REG), (m.TrainMode(False), O.TRAIN_NOREG), (m.TestMode(), O.TEST)]:
        args = (xs,) if ys is None else (xs, ys)
        l, g, gx = m.loss_and_gradient(icnf, mode, *args, theta, {}, want_dxs=True, eps=eps, tspan=icnf.tspan, **sol)
        rl, rg, rgx = O.loss_grad(om, omode, t64(xs), t64(theta), t64(eps), t64(ys), opts=opts, want_dxs=True)
        assert abs(l - float(rl)) <= RTOL * abs(float(rl)) + 1e-6
        assert norm_rel_err(g, rg.numpy()) < RTOL, (shape, mode, norm_rel_err(g, rg.numpy()))
        assert norm_rel_err(gx, rgx.numpy()) < RTOL
        np.testing.assert_allclose(g, rg.numpy(), rtol=5e-3, atol=2e-4 * float(rg.abs().max()))




def test_full_size_properties_config3(m):
    """BASELINE config 3 at its full batch (262 144, 16-D): size-independent properties instead of an
    oracle run -- (i) a linear field has a closed-form log-density, (ii) flowing forward then
    backward returns the input, (iii) sharding the batch does not change any sample's result."""
    import math
    import torch
    B, D = 262144, 16
    rng = np.random.default_rng(0)
    # (i) dz/dt = A z through the generic family (one Dense layer, identity)
    Amat = (0.15 * rng.standard_normal((D, D))).astype(np.float32)
    lin = m.ICNF(nvariables=D, naugments=0, autonomous=True, nn=m.Chain(m.Dense(D, D)))
    assert lin.kernel_family == "generic"
    theta = np.concatenate([Amat.flatten(order="F"), np.zeros(D, np.float32)])
    xs = rng.standard_normal((D, B)).astype(np.float32)
    logp, _ = m.inference(lin, m.TestMode(), xs, theta, {}, reltol=1e-6, abstol=1e-6)
    z1 = torch.matrix_exp(torch.tensor(Amat, dtype=torch.float64)).numpy() @ xs.astype(np.float64)
    want = -0.5 * D * math.log(2 * math.pi) - 0.5 * (z1 ** 2).sum(0) + float(np.trace(Amat.astype(np.float64)))
    np.testing.assert_allclose(logp, want, rtol=RTOL, atol=1e-4)
    # (ii) + (iii) on the config-3 network
    icnf = make_icnf(m, "config3_gmm16")
    om, th, _, _, _ = make_inputs(icnf, 1)
    u0 = np.zeros((D + 3, B), np.float32, order="F")
    u0[:D] = xs
    kw = dict(adaptive=False, dt=0.125)
    fwd = m.base_sol(icnf, m.TestMode(), u0, th, tspan=(0.0, 1.0), **kw)
    back = m.base_sol(icnf, m.TestMode(), np.asfortranarray(fwd), th, tspan=(1.0, 0.0), **kw)
    assert np.abs(back[:D] - u0[:D]).max() < 2e-4
    part = m.base_sol(icnf, m.TestMode(), np.asfortranarray(u0[:, 1000:3000]), th, tspan=(0.0, 1.0), **kw)
    np.testing.assert_allclose(part, fwd[:, 1000:3000], rtol=1e-6, atol=1e-6)
