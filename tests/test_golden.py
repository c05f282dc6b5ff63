"""Golden vectors (tests/golden/icnf_golden.npz, made by tests/golden/make_golden.py from the
float64 oracle): the oracle must keep reproducing them (CPU), and the CUDA path must match them
through the C ABI (GPU) without importing the oracle.  Fixed step dt = 1/8, tolerance 1e-4."""
import os

import numpy as np
import pytest

GOLD = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "icnf_golden.npz"))
CASES = {
    "usage_1d": dict(nvariables=1),
    "moons_2d": dict(nvariables=2, naugments=0),
    "cond_2d": dict(nvariables=2, naugments=1, nconditions=2, n_hidden=8),
    "gmm_16d": dict(nvariables=16, naugments=0),
}
ORACLE_KW = {
    "usage_1d": dict(nvars=1), "moons_2d": dict(nvars=2, naug=0),
    "cond_2d": dict(nvars=2, naug=1, ncond=2, hidden=(8, 8)), "gmm_16d": dict(nvars=16, naug=0),
}
RTOL = 1e-4


def _get(name, key):
    k = f"{name}/{key}"
    return GOLD[k] if k in GOLD.files else None


@pytest.mark.parametrize("name", list(CASES))
def test_oracle_reproduces_golden_vectors(name):
    import torch
    from oracle import icnf_oracle as O
    om = O.OracleICNF(**ORACLE_KW[name])
    t = lambda a: None if a is None else torch.tensor(a, dtype=torch.float64)
    theta, xs, eps, ys, u = (_get(name, k) for k in ("theta", "xs", "eps", "ys", "u"))
    opts = O.SolverOpts(adaptive=False, dt=0.125)
    for mode, tag in ((O.TEST, "test"), (O.TRAIN_REG, "train")):
        np.testing.assert_allclose(O.rhs_closed(om, mode, t(u), t(theta), 0.37, t(eps), t(ys)).numpy(), _get(name, f"rhs_{tag}"),
                                   rtol=1e-12, atol=1e-12)
        np.testing.assert_allclose(O.rhs_ad(om, mode, t(u), t(theta), 0.37, t(eps), t(ys)).detach().numpy(),
                                   _get(name, f"rhs_{tag}"), rtol=1e-10, atol=1e-11)
        logp, _ = O.inference(om, mode, t(xs), t(theta), t(eps), t(ys), opts=opts)
        np.testing.assert_allclose(logp.numpy(), _get(name, f"logp_{tag}"), rtol=1e-12, atol=1e-12)
    if name != "gmm_16d":      # (autograd through 16 pullbacks per stage is slow; the GPU test covers it)
        val, g, gx = O.loss_grad(om, O.TRAIN_REG, t(xs), t(theta), t(eps), t(ys), opts=opts, want_dxs=True)
        np.testing.assert_allclose(g.numpy(), _get(name, "dtheta_train"), rtol=1e-9, atol=1e-12)


@pytest.mark.gpu
@pytest.mark.parametrize("name", list(CASES))
def test_cuda_path_matches_golden_vectors(name):
    import cnf_b200 as m
    icnf = m.ICNF(**CASES[name])
    theta, xs, eps, ys, u = (_get(name, k) for k in ("theta", "xs", "eps", "ys", "u"))
    args = (xs,) if ys is None else (xs, ys)
    kw = dict(eps=eps, tspan=icnf.tspan, adaptive=False, dt=0.125)
    for mode, tag in ((m.TestMode(), "test"), (m.TrainMode(True), "train")):
        du = m.augmented_f(icnf, mode, u, theta, 0.37, eps=eps, ys=ys)
        np.testing.assert_allclose(du, _get(name, f"rhs_{tag}"), rtol=RTOL, atol=2e-5)
        logp, (E, n, A) = m.inference(icnf, mode, *args, theta, {}, **kw)
        np.testing.assert_allclose(logp, _get(name, f"logp_{tag}"), rtol=RTOL, atol=2e-5)
        np.testing.assert_allclose(np.stack([E, n, A]), _get(name, f"regs_{tag}"), rtol=RTOL, atol=2e-5)
    l, g, gx = m.loss_and_gradient(icnf, m.TrainMode(True), *args, theta, {}, want_dxs=True, **kw)
    assert abs(l - float(_get(name, "loss_train"))) <= RTOL * abs(float(_get(name, "loss_train")))
    ref = _get(name, "dtheta_train")
    assert np.linalg.norm(g - ref) / np.linalg.norm(ref) < RTOL
    refx = _get(name, "dxs_train")
    assert np.linalg.norm(gx - refx) / np.linalg.norm(refx) < RTOL
