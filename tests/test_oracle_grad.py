"""Oracle gradient (reverse AD through the discrete solve) vs float64 central
finite differences (SURVEY.md 8(c) check 5)."""
import numpy as np
import pytest
import torch

from oracle import icnf_oracle as O

F64 = torch.float64


@pytest.mark.parametrize("mode", [O.TRAIN_REG, O.TRAIN_NOREG, O.TEST])
@pytest.mark.parametrize("cond", [False, True])
def test_loss_grad_matches_finite_differences(mode, cond):
    m = O.OracleICNF(nvars=2, naug=1, ncond=2 if cond else 0, hidden=(5, 4))
    g = torch.Generator().manual_seed(1)
    theta = torch.tensor(O.init_params(m, 2, np.float64, bias_scale=0.2))
    xs = torch.randn(2, 5, dtype=F64, generator=g)
    eps = torch.randn(3, 5, dtype=F64, generator=g)
    ys = torch.randn(2, 5, dtype=F64, generator=g) if cond else None
    opts = O.SolverOpts(adaptive=False, dt=0.25)
    val, gth, gxs = O.loss_grad(m, mode, xs, theta, eps, ys, opts=opts, want_dxs=True)
    f = lambda th, x: float(O.loss(m, mode, x, th, eps, ys, opts=opts))
    assert float(val) == pytest.approx(f(theta, xs), rel=1e-12)
    h = 1e-6
    rng = np.random.default_rng(0)
    for i in rng.choice(theta.numel(), 12, replace=False):
        e = torch.zeros_like(theta)
        e[i] = h
        fd = (f(theta + e, xs) - f(theta - e, xs)) / (2 * h)
        assert float(gth[i]) == pytest.approx(fd, rel=2e-6, abs=1e-8)
    for i in range(xs.numel()):
        e = torch.zeros_like(xs).reshape(-1)
        e[i] = h
        e = e.reshape(xs.shape)
        fd = (f(theta, xs + e) - f(theta, xs - e)) / (2 * h)
        assert float(gxs.reshape(-1)[i]) == pytest.approx(fd, rel=2e-6, abs=1e-8)


def test_adaptive_grad_treats_step_sizes_as_constants():
    m = O.OracleICNF(nvars=1)
    g = torch.Generator().manual_seed(0)
    theta = torch.tensor(O.init_params(m, 0, np.float64, bias_scale=0.1))
    xs = torch.rand(1, 8, dtype=F64, generator=g)
    eps = torch.randn(3, 8, dtype=F64, generator=g)
    st = O.SolveStats()
    val, gth, _ = O.loss_grad(m, O.TRAIN_REG, xs, theta, eps, stats=st)
    assert st.naccept >= 1 and torch.isfinite(gth).all()
    # replay the accepted steps as a fixed schedule: same loss, same gradient
    def replay(th):
        f = lambda u, t: O.rhs_ad(m, O.TRAIN_REG, u, th, t, eps, create_graph=True)
        u = O.make_u0(m, xs)
        for t, dt in zip(st.ts, st.dts):
            u, _, _ = O.tsit5_step(f, u, f(u, t), t, dt)
        logp, (E, n, A) = O.readout(m, O.TRAIN_REG, u)
        return (-logp + m.lam1 * E + m.lam2 * n + m.lam3 * A).mean()
    th = theta.clone().requires_grad_(True)
    v2 = replay(th)
    (g2,) = torch.autograd.grad(v2, th)
    assert float(v2.detach()) == pytest.approx(float(val), rel=1e-12)
    assert torch.allclose(g2, gth, rtol=1e-9, atol=1e-12)
