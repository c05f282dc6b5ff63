"""GPU tests of the VCABM stepper (the reference's default ``alg``, icnf.jl:89) in the tiny family and on the narrow
single-launch path: the kernels follow
the oracle's restatement (oracle/icnf_oracle.py vcabm_solve: Hairer-Norsett-Wanner III.5 + Shampine-Gordon order
selection) operation for operation; fp32 rounding may move an accept/reject or an order decision, so results are
compared at tolerance level (as for adaptive Tsit5) and step counts within a small margin."""
import numpy as np
import pytest
import torch

from oracle import icnf_oracle as O
from tests.helpers import make_icnf, make_inputs, t64

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def m():
    import cnf_b200
    return cnf_b200


@pytest.mark.parametrize("shape", ["config2_moons", "config1_usage", "cond", "nvars3_default", "cond_generic"])
def test_vcabm_inference_matches_the_oracle(m, shape):
    if shape == "nvars3_default":       # ICNF(nvariables = 3) with every default: 8-32-32-7 softplus, the narrow single-launch path
        icnf = m.ICNF(nvariables=3)
    else:
        icnf = make_icnf(m, shape)
    assert icnf.solve_path(m.TestMode()) in ("tiny", "narrow")
    om, theta, xs, eps, ys = make_inputs(icnf, 500)
    theta = (1.5 * theta).astype(np.float32)
    args = (xs,) if ys is None else (xs, ys)
    tight = O.SolverOpts(reltol=1e-10, abstol=1e-10)
    for mode, omode in ((m.TestMode(), O.TEST), (m.TrainMode(True), O.TRAIN_REG)):
        exact, _ = O.inference(om, omode, t64(xs), t64(theta), t64(eps), t64(ys), opts=tight)
        for tol, margin in ((1e-4, 2e-3), (1e-6, 4e-5)):
            logp, (E, n, A) = m.inference(icnf, mode, *args, theta, {}, eps=eps, tspan=icnf.tspan, alg="VCABM", reltol=tol, abstol=tol)
            gs = icnf.last_stats
            st = O.SolveStats()
            rl, (rE, rn, rA) = O.inference(om, omode, t64(xs), t64(theta), t64(eps), t64(ys),
                                           opts=O.SolverOpts(alg="vcabm", reltol=tol, abstol=tol), stats=st)
            assert gs.status == 0 and gs.t_final == pytest.approx(1.0)
            assert gs.nf == 2 + 2 * (gs.naccept + gs.nreject)          # the final evaluation of a rejected attempt is speculative
            # step decisions are compared with the oracle run in float32 (the like-for-like: in float64 the order
            # selection takes other branches at 1e-6, e.g. 17 + 5 steps instead of 18 + 1)
            st32 = O.SolveStats()
            f32 = lambda a: None if a is None else torch.tensor(np.asarray(a), dtype=torch.float32)
            O.inference(om, omode, f32(xs), f32(theta), f32(eps), f32(ys), opts=O.SolverOpts(alg="vcabm", reltol=tol, abstol=tol), stats=st32)
            assert abs(gs.naccept - st32.naccept) <= 2 and abs(gs.nreject - st32.nreject) <= 3, (gs, st32.naccept, st32.nreject)
            np.testing.assert_allclose(logp, exact.numpy(), rtol=margin, atol=margin)
            np.testing.assert_allclose(logp, rl.numpy(), rtol=margin, atol=margin)
            np.testing.assert_allclose(E, rE.numpy(), rtol=margin, atol=margin)


def test_vcabm_generate_runs_backwards(m):
    icnf = make_icnf(m, "config2_moons")
    om, theta, xs, eps, ys = make_inputs(icnf, 300)
    z0 = np.random.default_rng(4).standard_normal((2, 300)).astype(np.float32)
    got = m.generate(icnf, m.TestMode(), theta, {}, 300, z0=z0, alg="VCABM", reltol=1e-6, abstol=1e-6)
    ref = O.generate(om, O.TEST, t64(z0), t64(theta), None, opts=O.SolverOpts(reltol=1e-10, abstol=1e-10)).numpy()
    np.testing.assert_allclose(got, ref, rtol=4e-5, atol=4e-5)


def test_vcabm_is_refused_where_it_is_not_served_and_gradients_use_tsit5(m):
    wide = m.ICNF(nvariables=64, naugments=0, nconditions=32, n_hidden=256)     # multi-launch generic path: Tsit5 only
    om, theta, xs, eps, ys = make_inputs(wide, 64)
    with pytest.raises(m.ICNFError):
        m.inference(wide, m.TestMode(), xs, ys, theta, {}, alg="VCABM")
    with pytest.raises(ValueError):
        m.ICNF(nvariables=2, sol_kwargs=dict(alg="Rodas5"))
    icnf = make_icnf(m, "config2_moons")
    om, theta, xs, eps, ys = make_inputs(icnf, 256)
    l1, g1 = m.loss_and_gradient(icnf, m.TrainMode(True), xs, theta, {}, eps=eps, tspan=icnf.tspan, alg="VCABM")
    l2, g2 = m.loss_and_gradient(icnf, m.TrainMode(True), xs, theta, {}, eps=eps, tspan=icnf.tspan, alg="Tsit5")
    assert l1 == l2 and np.array_equal(g1, g2)      # the reverse sweep differentiates Tsit5 steps whatever alg says
    lv = m.loss(icnf, m.TrainMode(True), xs, theta, {}, eps=eps, tspan=icnf.tspan, alg="VCABM")
    assert abs(lv - l2) < 2e-3 * abs(l2)
