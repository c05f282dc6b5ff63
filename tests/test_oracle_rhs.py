"""Self-validation of the oracle's right-hand side (SURVEY.md 8(c) checks 1-2)."""
import numpy as np
import pytest
import torch

from oracle import icnf_oracle as O

F64 = torch.float64


def _setup(model, B, seed=0):
    g = torch.Generator().manual_seed(seed)
    theta = torch.tensor(O.init_params(model, seed, np.float64, bias_scale=0.3))
    u = torch.randn(model.n_state, B, dtype=F64, generator=g)
    eps = torch.randn(model.d, B, dtype=F64, generator=g)
    ys = torch.randn(model.ncond, B, dtype=F64, generator=g) if model.ncond else None
    return theta, u, eps, ys


MODELS = [
    O.OracleICNF(nvars=1),
    O.OracleICNF(nvars=2, naug=0, hidden=(12, 12)),
    O.OracleICNF(nvars=2, ncond=2),
    O.OracleICNF(nvars=3, naug=1, autonomous=True, activation=O.ACT_TANH, hidden=(8,)),
    O.OracleICNF(nvars=2, naug=0, hidden=(7, 9, 5), activation=O.ACT_SIGMOID),
]


@pytest.mark.parametrize("model", MODELS)
@pytest.mark.parametrize("mode", [O.TEST, O.TRAIN_REG, O.TRAIN_NOREG])
def test_ad_and_closed_forms_agree(model, mode):
    theta, u, eps, ys = _setup(model, 6)
    a = O.rhs_ad(model, mode, u, theta, 0.37, eps, ys)
    c = O.rhs_closed(model, mode, u, theta, 0.37, eps, ys)
    assert a.shape == (model.n_state, 6)
    assert torch.allclose(a.detach(), c, rtol=1e-12, atol=1e-12)


@pytest.mark.parametrize("model", MODELS)
def test_exact_trace_and_vjp_against_full_jacobian(model):
    theta, u, eps, ys = _setup(model, 4, seed=1)
    params = O.unpack_params(theta, model.sizes)
    d = model.d
    for b in range(4):
        yb = ys[:, b:b + 1] if ys is not None else None
        f = lambda z: O.mlp(model, O.net_input(model, z[:, None], 0.2, yb), params)[:, 0]
        J = torch.autograd.functional.jacobian(f, u[:d, b])
        du_test = O.rhs_closed(model, O.TEST, u[:, b:b + 1], theta, 0.2, None, yb)
        assert torch.allclose(du_test[d, 0], -torch.trace(J), rtol=1e-11, atol=1e-12)
        du_tr = O.rhs_closed(model, O.TRAIN_REG, u[:, b:b + 1], theta, 0.2, eps[:, b:b + 1], yb)
        eJ = eps[:, b] @ J
        assert torch.allclose(du_tr[d, 0], -(eJ @ eps[:, b]), rtol=1e-11, atol=1e-12)
        assert torch.allclose(du_tr[d + 1, 0], torch.linalg.norm(f(u[:d, b])), rtol=1e-11)
        assert torch.allclose(du_tr[d + 2, 0], torch.linalg.norm(eJ), rtol=1e-11)


def test_regularisers_follow_mode_and_lambda_switches():
    m = O.OracleICNF(nvars=2, lam1=0.0)
    theta, u, eps, _ = _setup(m, 3)
    du = O.rhs_closed(m, O.TRAIN_REG, u, theta, 0.1, eps)
    assert torch.all(du[m.d + 1] == 0) and torch.all(du[m.d + 2] > 0)
    du = O.rhs_closed(m, O.TRAIN_NOREG, u, theta, 0.1, eps)
    assert torch.all(du[m.d + 1:] == 0)
    du = O.rhs_closed(m, O.TEST, u, theta, 0.1, eps)
    assert torch.all(du[m.d + 1:] == 0)
    msq = O.OracleICNF(nvars=2, reg_squared=True)
    a = O.rhs_closed(msq, O.TRAIN_REG, u, theta, 0.1, eps)
    b = O.rhs_closed(O.OracleICNF(nvars=2), O.TRAIN_REG, u, theta, 0.1, eps)
    assert torch.allclose(a[m.d + 1:], b[m.d + 1:] ** 2)


def test_hutchinson_mean_converges_to_exact_trace():
    m = O.OracleICNF(nvars=2, naug=1)
    theta, u, _, _ = _setup(m, 1)
    n = 20000
    ub = u.expand(-1, n).contiguous()
    g = torch.Generator().manual_seed(5)
    for eps in (torch.randn(m.d, n, dtype=F64, generator=g),
                torch.randint(0, 2, (m.d, n), generator=g).to(F64) * 2 - 1):
        est = O.rhs_closed(m, O.TRAIN_NOREG, ub, theta, 0.5, eps)[m.d].mean()
        exact = O.rhs_closed(m, O.TEST, u, theta, 0.5, None)[m.d, 0]
        assert abs(float(est - exact)) < 0.02 * max(1.0, abs(float(exact)))


def test_input_order_is_z_t_ys():
    # cond_layer.jl: outer layer appends t, inner appends ys -> [z; t; ys]
    m = O.OracleICNF(nvars=1, naug=0, ncond=1, hidden=(2,), activation=O.ACT_IDENTITY)
    # W1 picks columns: unit0 = t, unit1 = ys; W2 sums with weights (10, 100)
    W1 = np.array([[0.0, 1.0, 0.0], [0.0, 0.0, 1.0]])
    W2 = np.array([[10.0, 100.0]])
    theta = torch.tensor(np.concatenate([W1.flatten(order="F"), np.zeros(2), W2.flatten(order="F"), np.zeros(1)]))
    u = torch.zeros(4, 1, dtype=F64)
    ys = torch.tensor([[3.0]], dtype=F64)
    du = O.rhs_closed(m, O.TEST, u, theta, 0.5, None, ys)
    assert float(du[0, 0]) == pytest.approx(10 * 0.5 + 100 * 3.0)
