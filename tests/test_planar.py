"""PlanarLayer (src/layers/planar_layer.jl): f(x) = u * act(w'x + b) is served as Dense(n_in => 1, act) followed by a
bias-free Dense(1 => n_out); the host side re-orders the parameters (u, w, b) <-> [w; b; u; 0].  CPU: the mapping and its
adjoint; GPU: the reference's own smoke-test shape (smoke_tests.jl:30-45) against the oracle."""
import numpy as np
import pytest
import torch

from oracle import icnf_oracle as O


def planar_reference(u, w, b, x, act):
    return np.outer(u, act(w @ x + b))


def test_parameter_mapping_and_its_adjoint():
    import cnf_b200 as m
    pl = m.PlanarLayer(6, 5, "tanh")
    assert pl.n_params == 5 + 6 + 1
    rng = np.random.default_rng(0)
    ps = rng.standard_normal(pl.n_params).astype(np.float32)
    lib = pl.to_lib(ps)
    assert lib.shape == (6 * 1 + 1 + 1 * 5 + 5,)
    u, w, b = ps[:5], ps[5:11], ps[11]
    # the Dense pair with these parameters computes the planar layer
    om = O.OracleICNF(nvars=2, naug=3, hidden=(1,), activation=O.ACT_TANH)
    x = rng.standard_normal((6, 7))
    got = O.mlp(om, torch.tensor(x), O.unpack_params(torch.tensor(lib, dtype=torch.float64), (6, 1, 5))).numpy()
    ref = planar_reference(u.astype(np.float64), w.astype(np.float64), float(b), x, np.tanh)
    np.testing.assert_allclose(got, ref, rtol=1e-6, atol=1e-7)
    # adjoint: <to_lib(dp), g> = <dp, grad_from_lib(g)> for every dp, g
    dp = rng.standard_normal(pl.n_params).astype(np.float32)
    g = rng.standard_normal(lib.size).astype(np.float32)
    assert np.dot(pl.to_lib(dp), g) == pytest.approx(np.dot(dp, pl.grad_from_lib(g)), rel=1e-5)
    # torch and numpy agree; bias-free variant
    np.testing.assert_array_equal(pl.to_lib(torch.tensor(ps)).numpy(), lib)
    nb = m.PlanarLayer(6, 5, "tanh", use_bias=False)
    assert nb.n_params == 11 and nb.to_lib(ps[:11])[6] == 0.0 and nb.grad_from_lib(g).size == 11


@pytest.mark.gpu
@pytest.mark.parametrize("conditioned", [False, True])
def test_planar_icnf_matches_oracle(conditioned):
    import cnf_b200 as m
    from tests.helpers import t64, norm_rel_err
    nvars = 2
    ncond = nvars if conditioned else 0
    n_in = nvars * 2 + 2 + ncond           # D' = 2 nvars + 1 (default naugments = nvars + 1), + time (+ conditions)
    icnf = m.ICNF(nvariables=nvars, nconditions=ncond, nn=m.Chain(m.PlanarLayer(n_in, nvars * 2 + 1, "tanh")))
    assert icnf.n_params == (nvars * 2 + 1) + n_in + 1
    om = O.OracleICNF(nvars=nvars, naug=nvars + 1, ncond=ncond, hidden=(1,), activation=O.ACT_TANH,
                      lam1=icnf.lambda1, lam2=icnf.lambda2, lam3=icnf.lambda3, tspan=icnf.tspan)
    rng = np.random.default_rng(1)
    ps, _ = m.setup(rng, icnf)
    ps = (ps + 0.3 * rng.standard_normal(ps.size)).astype(np.float32)
    theta = icnf.planar.to_lib(ps)
    B = 200
    xs = rng.standard_normal((nvars, B)).astype(np.float32)
    ys = rng.standard_normal((ncond, B)).astype(np.float32) if ncond else None
    eps = rng.standard_normal((om.d, B)).astype(np.float32)
    args = (xs,) if ys is None else (xs, ys)
    for mode, omode in ((m.TestMode(), O.TEST), (m.TrainMode(True), O.TRAIN_REG)):
        logp, (E, n, A) = m.inference(icnf, mode, *args, ps, {}, eps=eps, tspan=icnf.tspan)
        rl, (rE, rn, rA) = O.inference(om, omode, t64(xs), t64(theta), t64(eps), t64(ys))
        np.testing.assert_allclose(logp, rl.numpy(), rtol=1e-4, atol=2e-5)
        np.testing.assert_allclose(A, rA.numpy(), rtol=1e-4, atol=2e-5)
    l, g = m.loss_and_gradient(icnf, m.TrainMode(True), *args, ps, {}, eps=eps, tspan=icnf.tspan, adaptive=False, dt=0.25)
    rl, rg, _ = O.loss_grad(om, O.TRAIN_REG, t64(xs), t64(theta), t64(eps), t64(ys), opts=O.SolverOpts(adaptive=False, dt=0.25))
    assert g.shape == (icnf.n_params,)
    assert abs(l - float(rl)) <= 1e-4 * abs(float(rl))
    assert norm_rel_err(g, icnf.planar.grad_from_lib(rg.numpy())) < 2e-4
    # the output bias of the Dense pair is not a parameter of the planar layer: its gradient never reaches the caller
    zs = m.generate(icnf, m.TestMode(), *(() if ys is None else (ys,)), ps, {}, B, z0=eps)
    assert zs.shape == (nvars, B) and np.isfinite(zs).all()
