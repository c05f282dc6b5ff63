"""The gradient semantics gap: the product differentiates the DISCRETE Tsit5
solve (discretise-then-optimise); the reference returns a tolerance-1e-4 CONTINUOUS adjoint
(/root/reference/src/core/icnf.jl:90-99).  These CPU tests pin the oracle's continuous adjoint and
state how far apart the two gradients are, so the deviation is a number (quoted in DESIGN.md)."""
import numpy as np
import torch

from oracle import icnf_oracle as O


def _case(seed=0, B=24, scale=1.0):
    om = O.OracleICNF(nvars=2, naug=1, hidden=(10, 10))
    rng = np.random.default_rng(seed)
    theta = scale * torch.tensor(O.init_params(om, seed + 1, np.float64, bias_scale=0.3))
    xs = torch.tensor(rng.standard_normal((2, B)))
    eps = torch.tensor(rng.standard_normal((om.d, B)))
    return om, theta, xs, eps


def _rel(a, b):
    return float(torch.linalg.norm(a - b) / torch.linalg.norm(b))


def test_continuous_adjoint_matches_discrete_gradient_at_tight_tolerance():
    om, theta, xs, eps = _case()
    tight = O.SolverOpts(reltol=1e-10, abstol=1e-10)
    lc, gc, gxc = O.loss_grad_continuous(om, O.TRAIN_REG, xs, theta, eps, opts=tight, want_dxs=True)
    ld, gd, gxd = O.loss_grad(om, O.TRAIN_REG, xs, theta, eps, opts=O.SolverOpts(reltol=1e-9, abstol=1e-9), want_dxs=True)
    assert abs(float(lc) - float(ld)) < 1e-8 * abs(float(ld))
    assert _rel(gc, gd) < 1e-6, _rel(gc, gd)
    assert _rel(gxc, gxd) < 1e-6, _rel(gxc, gxd)


def test_continuous_adjoint_exact_trace_mode():
    om, theta, xs, eps = _case(3, B=8)
    tight = O.SolverOpts(reltol=1e-10, abstol=1e-10)
    lc, gc, _ = O.loss_grad_continuous(om, O.TEST, xs, theta, None, opts=tight)
    ld, gd, _ = O.loss_grad(om, O.TEST, xs, theta, None, opts=O.SolverOpts(reltol=1e-9, abstol=1e-9))
    assert _rel(gc, gd) < 1e-6, _rel(gc, gd)


def test_gap_between_discrete_and_continuous_gradients_at_reference_tolerance(capsys):
    """At reltol = abstol = 1e-4 (icnf.jl:87-88) both gradients carry O(tolerance) error; they sit at the
    same distance from the exact gradient, and from each other.  The weights are scaled x3 so that the
    flow is strong enough for the controller to work (13 accepted steps); with the plain glorot draw the
    solve takes 3 steps and all three gradients agree to 1e-8."""
    om, theta, xs, eps = _case(scale=3.0)
    tight = O.SolverOpts(reltol=1e-10, abstol=1e-10)
    ref_tol = O.SolverOpts(reltol=1e-4, abstol=1e-4)
    _, g_exact, _ = O.loss_grad_continuous(om, O.TRAIN_REG, xs, theta, eps, opts=tight)
    _, g_disc, _ = O.loss_grad(om, O.TRAIN_REG, xs, theta, eps, opts=ref_tol)
    _, g_cont, _ = O.loss_grad_continuous(om, O.TRAIN_REG, xs, theta, eps, opts=ref_tol)
    gap_disc, gap_cont, gap_dc = _rel(g_disc, g_exact), _rel(g_cont, g_exact), _rel(g_disc, g_cont)
    with capsys.disabled():
        print(f"\n[gradient gap at tol 1e-4] |disc - exact| = {gap_disc:.2e}  |cont - exact| = {gap_cont:.2e}  "
              f"|disc - cont| = {gap_dc:.2e}  (relative L2)")
    assert gap_disc < 2e-3 and gap_cont < 2e-3 and gap_dc < 2e-3
