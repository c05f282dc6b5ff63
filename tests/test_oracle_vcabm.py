"""The oracle's restatement of VCABM (the reference's default ``alg``, icnf.jl:89): variable-step variable-order
Adams-Bashforth-Moulton PECE after Hairer-Norsett-Wanner III.5.  Validated against closed forms, scipy and the
oracle's own Tsit5 on the augmented ICNF system (the third-party solver itself cannot run here: parity unpinned)."""
import math

import numpy as np
import torch
from scipy.integrate import solve_ivp

from oracle import icnf_oracle as O


def test_constant_step_coefficients_are_the_adams_bashforth_ones():
    # equal steps: beta_j = 1 and g_j = gamma_j = 1, 1/2, 5/12, 3/8, 251/720 (HNW III.1)
    beta, g = O.vcabm_coefficients([0.1] * 8, 4, 5)
    np.testing.assert_allclose(beta, [1.0] * 5, rtol=1e-14)
    np.testing.assert_allclose(g, [1.0, 1 / 2, 5 / 12, 3 / 8, 251 / 720], rtol=1e-13)
    # gamma*_j = gamma_j - gamma_{j-1} (HNW III.1 (1.10))
    gam = [1.0, 1 / 2, 5 / 12, 3 / 8, 251 / 720, 95 / 288]
    for j in range(1, 6):
        assert O.GAMMA_STAR[j] == np.float64(gam[j] - gam[j - 1]) or abs(O.GAMMA_STAR[j] - (gam[j] - gam[j - 1])) < 1e-15


def test_variable_step_predictor_is_exact_for_polynomials():
    # the order-k predictor integrates polynomials of degree < k exactly on ANY grid: one PECE step of y' = 3 t^2
    # from an irregular history reproduces t^3
    f = lambda u, t: torch.full_like(u, 3.0 * t * t)
    for tol in (1e-6,):
        st = O.SolveStats()
        y = O.vcabm_solve(f, torch.zeros(1, 1, dtype=torch.float64), 0.0, 2.0, O.SolverOpts(alg="vcabm", reltol=tol, abstol=tol), st)
        assert abs(float(y) - 8.0) < 1e-4
        assert len(set(np.round(st.dts, 12))) > 3      # the grid really is irregular


def test_error_tracks_the_tolerance_and_order_rises():
    f = lambda u, t: -u + math.sin(t)
    ref = solve_ivp(lambda t, y: -y + np.sin(t), (0, 5), [1.0], rtol=1e-13, atol=1e-13).y[0, -1]
    errs, nfs = [], []
    for tol in (1e-4, 1e-6, 1e-8):
        st = O.SolveStats()
        y = O.vcabm_solve(f, torch.ones(1, 1, dtype=torch.float64), 0.0, 5.0, O.SolverOpts(alg="vcabm", reltol=tol, abstol=tol), st)
        errs.append(abs(float(y) - ref)); nfs.append(st.nf)
        assert st.nf == 2 + 2 * st.naccept + st.nreject      # PECE: two evaluations per accepted step, one per rejected one
        assert errs[-1] < 20 * tol
    assert errs[2] < errs[1] < errs[0]
    # a multistep method of rising order: 100x tighter tolerance costs far less than 100^(1/2) more evaluations
    assert nfs[2] < 2.5 * nfs[0]


def test_icnf_solve_agrees_with_tsit5_within_tolerance_in_both_directions():
    om = O.OracleICNF(nvars=2, naug=0)
    rng = np.random.default_rng(0)
    theta = torch.tensor(2.0 * O.init_params(om, 1, np.float64, bias_scale=0.3))
    xs = torch.tensor(rng.standard_normal((2, 64)))
    eps = torch.tensor(rng.standard_normal((2, 64)))
    tight = O.SolverOpts(reltol=1e-10, abstol=1e-10)
    for mode in (O.TEST, O.TRAIN_REG):
        ref, _ = O.inference(om, mode, xs, theta, eps, opts=tight)
        st = O.SolveStats()
        got, _ = O.inference(om, mode, xs, theta, eps, opts=O.SolverOpts(alg="vcabm"), stats=st)
        assert st.naccept > 5
        np.testing.assert_allclose(got.numpy(), ref.numpy(), rtol=2e-3, atol=2e-3)
        got6, _ = O.inference(om, mode, xs, theta, eps, opts=O.SolverOpts(alg="vcabm", reltol=1e-7, abstol=1e-7))
        np.testing.assert_allclose(got6.numpy(), ref.numpy(), rtol=2e-6, atol=2e-6)
    z0 = torch.tensor(rng.standard_normal((2, 64)))
    a = O.generate(om, O.TEST, z0, theta, None, opts=O.SolverOpts(alg="vcabm", reltol=1e-7, abstol=1e-7))
    b = O.generate(om, O.TEST, z0, theta, None, opts=tight)
    np.testing.assert_allclose(a.numpy(), b.numpy(), rtol=1e-5, atol=1e-5)
