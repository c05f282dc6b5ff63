"""GPU tests of the callers on either side of the hot path (SURVEY 8(f) n2 and the adapters):
the MLJ-style fit loop with the host and the device optimiser, and the Distributions adapter."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def m():
    import cnf_b200
    return cnf_b200


def test_device_adam_matches_host_optimiser_chain(m):
    from cnf_b200.mlj import Adam, WeightDecay, _DeviceOptimiserChain, _OptimiserChain
    rng = np.random.default_rng(0)
    theta0 = rng.standard_normal(403).astype(np.float32)
    host = _OptimiserChain(WeightDecay(), Adam(), theta0.size)
    devopt = _DeviceOptimiserChain(WeightDecay(), Adam(), theta0, 0)
    th = theta0.copy()
    for step in range(5):
        g = rng.standard_normal(403).astype(np.float32)
        th = host.step(th, g)
        devopt.step(torch.tensor(g, device="cuda"))
    np.testing.assert_allclose(devopt.theta.cpu().numpy(), th, rtol=2e-5, atol=1e-6)


@pytest.mark.parametrize("device_optimiser", [False, True])
def test_fit_reduces_the_loss_and_transform_returns_densities(m, device_optimiser):
    # regression_tests.jl in miniature: Beta(2, 4) data, default ICNF(nvariables = 1)
    rng = np.random.default_rng(1)
    X = rng.beta(2.0, 4.0, size=(512, 1)).astype(np.float32)
    icnf = m.ICNF(nvariables=1, rng=0)
    losses = []
    model = m.ICNFModel(icnf=icnf, batchsize=256, epochs=15, rng=0, device_optimiser=device_optimiser,
                        callback=lambda it, l: losses.append(l) and False)
    model.fit(X)
    assert model.report["iterations"] == 30
    assert np.mean(losses[-4:]) < np.mean(losses[:4])
    px = model.transform(X)["px"]
    assert px.shape == (512,) and np.isfinite(px).all() and (px > 0).all()
    fp = model.fitted_params()
    assert fp["learned_parameters"].shape == (403,)
    d = m.ICNFDist(icnf, m.TestMode(), *model.fitresult)
    assert d.rand(7).shape == (1, 7)


def test_conditional_model_and_dist(m):
    rng = np.random.default_rng(2)
    X = rng.standard_normal((128, 2)).astype(np.float32)
    Y = rng.standard_normal((128, 2)).astype(np.float32)
    icnf = m.ICNF(nvariables=2, naugments=1, nconditions=2, n_hidden=8, rng=0)
    model = m.CondICNFModel(icnf=icnf, batchsize=64, epochs=2, rng=0)
    model.fit(X, Y)
    px = model.transform(X, Y)["px"]
    assert px.shape == (128,) and np.isfinite(px).all()
    d = m.CondICNFDist(icnf, m.TestMode(), Y.T[:, :5].copy(), *model.fitresult)
    assert d.rand(5).shape == (2, 5)
    assert d.logpdf(X.T[:, :5].copy()).shape == (5,)
