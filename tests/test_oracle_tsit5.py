"""Tsit5 tableau identities, convergence order, and closed-form flow checks
(SURVEY.md 8(c) checks 3-4)."""
import math

import numpy as np
import pytest
import torch

from oracle import icnf_oracle as O

F64 = torch.float64


def test_tableau_identities():
    for i in range(1, 7):
        assert sum(O.A[i]) == pytest.approx(O.C[i], abs=1e-14)
    b = list(O.A[6]) + [0.0]
    assert sum(b) == pytest.approx(1.0, abs=1e-14)
    for k in range(1, 5):
        assert sum(bi * c ** k for bi, c in zip(b, O.C)) == pytest.approx(1.0 / (k + 1), abs=1e-13)
    assert sum(O.BT) == pytest.approx(0.0, abs=1e-14)
    bh = [bi - bt for bi, bt in zip(b, O.BT)]     # embedded 4th-order weights
    for k in range(1, 4):
        assert sum(bi * c ** k for bi, c in zip(bh, O.C)) == pytest.approx(1.0 / (k + 1), abs=1e-12)
    assert abs(sum(bi * c ** 4 for bi, c in zip(bh, O.C)) - 0.2) > 1e-6


def test_fifth_order_convergence():
    f = lambda u, t: torch.stack([u[1] * (1 + 0.3 * math.sin(t)), -u[0]])
    u0 = torch.tensor([[1.0], [0.0]], dtype=F64)
    ref = O.tsit5_solve(f, u0, 0.0, 2.0, O.SolverOpts(adaptive=False, dt=2.0 / 512))
    errs = []
    for n in (8, 16, 32):
        u = O.tsit5_solve(f, u0, 0.0, 2.0, O.SolverOpts(adaptive=False, dt=2.0 / n))
        errs.append(float((u - ref).abs().max()))
    o1 = math.log2(errs[0] / errs[1])
    o2 = math.log2(errs[1] / errs[2])
    assert 4.5 < o1 < 6.5 and 4.5 < o2 < 6.5, (errs, o1, o2)


def test_fixed_step_clips_last_step_and_counts_rhs_calls():
    f = lambda u, t: -u
    st = O.SolveStats()
    u = O.tsit5_solve(f, torch.ones(1, 1, dtype=F64), 0.0, 1.0, O.SolverOpts(adaptive=False, dt=0.3), st)
    assert st.naccept == 4 and st.nf == 1 + 6 * 4
    assert sum(st.dts) == pytest.approx(1.0)
    assert float(u) == pytest.approx(math.exp(-1.0), rel=1e-5)


def test_adaptive_meets_tolerance_and_runs_backwards():
    f = lambda u, t: torch.stack([u[1], -u[0]])
    u0 = torch.tensor([[1.0], [0.0]], dtype=F64)
    st = O.SolveStats()
    u = O.tsit5_solve(f, u0, 0.0, 3.0, O.SolverOpts(reltol=1e-6, abstol=1e-6), st)
    assert float(u[0]) == pytest.approx(math.cos(3.0), abs=2e-5)
    assert float(u[1]) == pytest.approx(-math.sin(3.0), abs=2e-5)
    assert st.nf == 2 + 6 * (st.naccept + st.nreject)
    back = O.tsit5_solve(f, u, 3.0, 0.0, O.SolverOpts(reltol=1e-6, abstol=1e-6))
    assert torch.allclose(back, u0, atol=5e-5)


def _linear_model(Amat):
    """An ICNF whose network is exactly z -> A z (identity activation, one layer)."""
    d = Amat.shape[0]
    m = O.OracleICNF(nvars=d, naug=0, hidden=(), autonomous=True, activation=O.ACT_IDENTITY)
    theta = torch.tensor(np.concatenate([Amat.flatten(order="F"), np.zeros(d)]), dtype=F64)
    return m, theta


def test_linear_field_has_closed_form_logdensity():
    # dz/dt = A z  =>  z(1) = expm(A) x,  dlogp(1) = -tr(A)
    rng = np.random.default_rng(0)
    Amat = 0.5 * rng.standard_normal((3, 3))
    m, theta = _linear_model(Amat)
    xs = torch.tensor(rng.standard_normal((3, 5)))
    eps = torch.tensor(rng.standard_normal((3, 5)))
    opts = O.SolverOpts(reltol=1e-9, abstol=1e-9)
    expA = torch.matrix_exp(torch.tensor(Amat))
    z1 = expA @ xs
    want = -0.5 * 3 * math.log(2 * math.pi) - 0.5 * (z1 ** 2).sum(0) + np.trace(Amat)
    for closed in (True, False):
        logp, (E, n, Adot) = O.inference(m, O.TEST, xs, theta, eps, opts=opts, closed=closed)
        assert torch.allclose(logp.detach(), want, rtol=1e-7, atol=1e-7)
        assert torch.all(E == 0) and torch.all(n == 0) and torch.all(Adot == 0)
    # Hutchinson with a linear field: -eps' A eps integrated over t in [0,1]
    logp_h, _ = O.inference(m, O.TRAIN_NOREG, xs, theta, eps, opts=opts)
    quad = torch.einsum("ib,ij,jb->b", eps, torch.tensor(Amat), eps)
    want_h = -0.5 * 3 * math.log(2 * math.pi) - 0.5 * (z1 ** 2).sum(0) + quad
    assert torch.allclose(logp_h, want_h, rtol=1e-7, atol=1e-7)


def test_generate_inverts_inference_flow():
    m = O.OracleICNF(nvars=2, naug=1)
    theta = torch.tensor(O.init_params(m, 3, np.float64, bias_scale=0.2))
    g = torch.Generator().manual_seed(0)
    xs = torch.randn(2, 6, dtype=F64, generator=g)
    eps = torch.randn(3, 6, dtype=F64, generator=g)
    opts = O.SolverOpts(reltol=1e-9, abstol=1e-9)
    fsol = O.solve(m, O.TEST, O.make_u0(m, xs), theta, eps, opts=opts)
    back = O.generate(m, O.TEST, fsol[:3], theta, eps, opts=opts)
    assert torch.allclose(back, xs, atol=1e-6)


def test_readout_and_loss_definitions():
    m = O.OracleICNF(nvars=1)     # naug = 2, lam = 0.01 each
    fsol = torch.tensor([[0.5], [1.0], [-2.0], [0.7], [0.3], [0.4]], dtype=F64)
    logp, (E, n, Adot) = O.readout(m, O.TRAIN_REG, fsol)
    assert float(logp) == pytest.approx(-1.5 * math.log(2 * math.pi) - 0.5 * 5.25 - 0.7)
    assert float(E) == 0.3 and float(n) == 0.4
    assert float(Adot) == pytest.approx(math.sqrt(5.0))
    _, (_, _, A0) = O.readout(m, O.TEST, fsol)
    assert float(A0) == 0.0
    assert O.steer_t1(m, O.TRAIN_REG, 0.5) == pytest.approx(1.5)
    assert O.steer_t1(m, O.TEST, 0.5) == 1.0


@pytest.mark.parametrize("mode", [O.TEST, O.TRAIN_REG])
def test_solve_agrees_with_an_independent_integrator(mode):
    """The oracle's own Tsit5 + controller against scipy's DOP853 (a different method, a third-party
    implementation) on the full augmented ODE of a nonlinear flow: both must converge to the same state."""
    from scipy.integrate import solve_ivp
    m = O.OracleICNF(nvars=1, naug=2)                       # config 1: 4-16-16-3 softplus
    theta = torch.tensor(1.5 * O.init_params(m, seed=4, dtype=np.float64, bias_scale=0.1), dtype=F64)
    B = 5
    rng = np.random.default_rng(9)
    xs = torch.tensor(rng.standard_normal((1, B)), dtype=F64)
    eps = None if mode == O.TEST else torch.tensor(rng.standard_normal((m.d, B)), dtype=F64)
    u0 = O.make_u0(m, xs)
    mine = O.solve(m, mode, u0, theta, eps, opts=O.SolverOpts(reltol=1e-10, abstol=1e-10))

    def f(t, y):
        u = torch.tensor(y.reshape(m.n_state, B), dtype=F64)
        return O.rhs_closed(m, mode, u, theta, t, eps, None).numpy().reshape(-1)

    ref = solve_ivp(f, (0.0, 1.0), u0.numpy().reshape(-1), method="DOP853", rtol=1e-11, atol=1e-12)
    assert ref.success
    np.testing.assert_allclose(mine.numpy().reshape(-1), ref.y[:, -1], rtol=1e-7, atol=1e-8)
    # and the default tolerance (1e-4, what the reference and the kernels run at) stays within it
    loose = O.solve(m, mode, u0, theta, eps, opts=O.SolverOpts())
    np.testing.assert_allclose(loose.numpy().reshape(-1), ref.y[:, -1], rtol=2e-3, atol=2e-4)
