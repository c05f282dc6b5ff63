"""Host optimiser chain of the MLJ adapter (OptimiserChain(WeightDecay, Adam), mlj_ext/core_icnf.jl:17-24)
against torch.optim.Adam with L2 weight decay -- an independent implementation of the same update rule."""
import numpy as np
import torch


def test_optimiser_chain_matches_torch_adam():
    from cnf_b200.mlj import Adam, WeightDecay, _OptimiserChain
    rng = np.random.default_rng(0)
    n = 57
    theta0 = rng.standard_normal(n).astype(np.float32)
    grads = [rng.standard_normal(n).astype(np.float32) * (1.0 + k) for k in range(25)]
    opt = _OptimiserChain(WeightDecay(), Adam(), n)
    th = theta0.copy()
    for g in grads:
        th = opt.step(th, g)
    p = torch.nn.Parameter(torch.tensor(theta0, dtype=torch.float64))
    ref = torch.optim.Adam([p], lr=Adam().eta, betas=Adam().beta, eps=Adam().epsilon, weight_decay=WeightDecay().lam)
    for g in grads:
        p.grad = torch.tensor(g, dtype=torch.float64)
        ref.step()
    np.testing.assert_allclose(th, p.detach().numpy(), rtol=2e-5, atol=2e-6)
