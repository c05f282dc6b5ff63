"""CPU restatement of ContinuousNormalizingFlows.jl's batched augmented-ODE path.

TEST INFRASTRUCTURE ONLY.  Only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import this
module.  The product (``continuousnormalizingflows.jl_b200``) never does: it
fails loudly when its CUDA library is missing.

PARITY UNPINNED.  The reference is pure Julia; Julia is not installed in the
build container or on the GPU box, the reference's tests hold no numeric golden
vectors (every assertion is ``!isnothing`` / ``@test true``,
/root/reference/test/ci_tests/smoke_tests.jl:69-156, regression_tests.jl:28),
and the stepper/AD arithmetic lives in third-party packages that are not
vendored (Lux 1.x, NNlib 0.9, OrdinaryDiffEq*, SciMLSensitivity 7, Zygote 0.7;
bounds only in /root/reference/Project.toml:35-63, no Manifest).  This file is
therefore a from-scratch restatement, validated by its own known-answer tests
(tests/test_oracle_*.py): closed-form log-density of a linear field, autograd
VJP vs full jacobian, Hutchinson mean -> exact trace, Tsit5 order of
convergence and tableau identities, the augmented solve vs scipy's DOP853,
gradients vs float64 finite differences.

Two independent statements of the RHS are kept on purpose:
  * ``rhs_ad``      mirrors HOW the reference computes it: network forward, then
                    a reverse-mode vector-Jacobian product (torch.autograd plays
                    Zygote's role) or D' one-hot pullbacks for the exact trace
                    (src/core/utils.jl:35-54, :150-159).
  * ``rhs_closed``  the closed forms of SURVEY.md Appendix B written as dense
                    matrix products; fast, no autograd; this is the CPU baseline
                    that bench.py times.
The CUDA kernels implement neither by translation; tests compare them with both.

Shapes follow the reference: matrices are ``R x B`` (one column per sample).

Reference map (all under /root/reference/src):
  core/icnf.jl:143-145     two extra state rows (E, n) after the logdet row
  core/icnf.jl:147-161     time concatenation (CondLayer), skipped if autonomous
  layers/cond_layer.jl     network input order [z; t; ys]
  core/icnf.jl:184-251     reg_z / reg_j: UN-squared column norms, TrainMode{true} only
  core/icnf.jl:297-316     TestMode RHS: du = [zdot; -tr J; 0; 0]
  core/icnf.jl:517-536     TrainMode RHS: du = [zdot; -sum(eJ .* e); |zdot|; |eJ|]
  core/base_icnf.jl:247-296 u0 = [xs; 0], eps drawn once per solve
  core/base_icnf.jl:158-172 readout: logp = logpdf(N(0,I), z) - dlogp
  core/base_icnf.jl:106-132 A = |z_aug| at t1 (TrainMode{true}, naug != 0, l3 != 0)
  core/base_icnf.jl:351-404, :185-194  generate: reversed span, keep first nvars rows
  core/base_icnf.jl:23-43  STEER end-time perturbation (TrainMode{true})
  core/icnf.jl:628-649     loss = mean(-logp + l1 E + l2 n + l3 A)
"""

from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Callable, List, Optional, Sequence, Tuple

import numpy as np
import torch

# --------------------------------------------------------------------------
# modes (src/core/types.jl:1-7)
TEST = 0          # TestMode: exact trace, regularisers zero
TRAIN_REG = 1     # TrainMode{true}: Hutchinson + regularisers
TRAIN_NOREG = 2   # TrainMode{false}: Hutchinson, regularisers zero

ACT_SOFTPLUS, ACT_TANH, ACT_SIGMOID, ACT_IDENTITY = 0, 1, 2, 3


@dataclass
class OracleICNF:
    """Plain-data mirror of the ``ICNF`` struct (src/core/icnf.jl:16-141)."""

    nvars: int = 1
    naug: Optional[int] = None          # default nvars + 1 (icnf.jl:62)
    ncond: int = 0
    autonomous: bool = False
    hidden: Optional[Sequence[int]] = None   # default (4 n_in, 4 n_in) (icnf.jl:66-71)
    activation: int = ACT_SOFTPLUS
    lam1: float = 0.01
    lam2: float = 0.01
    lam3: float = 0.01
    reg_squared: bool = False           # reference uses un-squared norms
    tspan: Tuple[float, float] = (0.0, 1.0)
    steer_rate: float = 0.1

    def __post_init__(self):
        if self.naug is None:
            self.naug = self.nvars + 1
        if self.hidden is None:
            self.hidden = (4 * self.n_in, 4 * self.n_in)
        self.hidden = tuple(int(h) for h in self.hidden)

    @property
    def d(self) -> int:                 # D' = nvars + naug
        return self.nvars + self.naug

    @property
    def n_in(self) -> int:              # icnf.jl:64
        return self.d + (0 if self.autonomous else 1) + self.ncond

    @property
    def sizes(self) -> Tuple[int, ...]:
        return (self.n_in,) + tuple(self.hidden) + (self.d,)

    @property
    def n_state(self) -> int:           # S = D' + 3
        return self.d + 3

    @property
    def n_params(self) -> int:
        s = self.sizes
        return sum(s[l] * s[l + 1] + s[l + 1] for l in range(len(s) - 1))


# --------------------------------------------------------------------------
# parameters: flat vector in ComponentArray order [vec(W1); b1; vec(W2); b2; ...]
# with W_l column-major n_l x n_{l-1}  (src/exts/mlj_ext/core_icnf.jl:35)

def unpack_params(theta: torch.Tensor, sizes: Sequence[int]) -> List[Tuple[torch.Tensor, torch.Tensor]]:
    out, off = [], 0
    for l in range(len(sizes) - 1):
        nin, nout = sizes[l], sizes[l + 1]
        W = theta[off: off + nin * nout].reshape(nin, nout).transpose(0, 1)  # column-major -> (nout, nin)
        off += nin * nout
        b = theta[off: off + nout]
        off += nout
        out.append((W, b))
    assert off == theta.numel(), (off, theta.numel())
    return out


def init_params(model: OracleICNF, seed: int = 0, dtype=np.float32, bias_scale: float = 0.0) -> np.ndarray:
    """Lux Dense default init: glorot_uniform weights, zero bias [3P].  A
    non-zero ``bias_scale`` makes tests exercise the bias path."""
    rng = np.random.default_rng(seed)
    s = model.sizes
    parts = []
    for l in range(len(s) - 1):
        nin, nout = s[l], s[l + 1]
        lim = math.sqrt(6.0 / (nin + nout))
        parts.append(rng.uniform(-lim, lim, size=nin * nout))
        parts.append(bias_scale * rng.standard_normal(nout))
    return np.concatenate(parts).astype(dtype)


# --------------------------------------------------------------------------
# activations

def act(x: torch.Tensor, kind: int) -> torch.Tensor:
    if kind == ACT_SOFTPLUS:
        # NNlib.softplus(x) = log1p(exp(-|x|)) + relu(x) [3P]; logaddexp(x, 0)
        # is the same function and is smooth under autograd at x == 0.
        return torch.logaddexp(x, torch.zeros((), dtype=x.dtype))
    if kind == ACT_TANH:
        return torch.tanh(x)
    if kind == ACT_SIGMOID:
        return torch.sigmoid(x)
    if kind == ACT_IDENTITY:
        return x
    raise ValueError(kind)


def act_d1(x: torch.Tensor, kind: int) -> torch.Tensor:
    if kind == ACT_SOFTPLUS:
        return torch.sigmoid(x)
    if kind == ACT_TANH:
        return 1.0 - torch.tanh(x) ** 2
    if kind == ACT_SIGMOID:
        s = torch.sigmoid(x)
        return s * (1.0 - s)
    if kind == ACT_IDENTITY:
        return torch.ones_like(x)
    raise ValueError(kind)


# --------------------------------------------------------------------------
# network input [z; t; ys]  (cond_layer.jl, icnf.jl:147-161, base_icnf.jl:49-60)

def net_input(model: OracleICNF, z: torch.Tensor, t, ys: Optional[torch.Tensor]) -> torch.Tensor:
    parts = [z]
    if not model.autonomous:
        parts.append(torch.full((1, z.shape[1]), float(t), dtype=z.dtype))
    if model.ncond:
        assert ys is not None and ys.shape == (model.ncond, z.shape[1])
        parts.append(ys.to(z.dtype))
    return torch.cat(parts, dim=0)


def mlp(model: OracleICNF, h0: torch.Tensor, params) -> torch.Tensor:
    h = h0
    L = len(params)
    for l, (W, b) in enumerate(params):
        a = W @ h + b[:, None]
        h = act(a, model.activation) if l < L - 1 else a
    return h


def _col_norm(x: torch.Tensor, squared: bool) -> torch.Tensor:
    s = (x * x).sum(dim=0)
    if squared:
        return s
    # d|x|/dx at 0 := 0, matching ChainRules' rule for norm [3P]
    safe = torch.where(s > 0, s, torch.ones_like(s))
    return torch.where(s > 0, torch.sqrt(safe), torch.zeros_like(s))


# --------------------------------------------------------------------------
# RHS, statement 1: the way the reference computes it (forward + reverse AD)

def rhs_ad(model: OracleICNF, mode: int, u: torch.Tensor, theta: torch.Tensor, t,
           eps: Optional[torch.Tensor], ys: Optional[torch.Tensor] = None,
           create_graph: bool = False) -> torch.Tensor:
    d = model.d
    params = unpack_params(theta, model.sizes)
    z = u[:d, :]
    if not z.requires_grad:
        z = z.detach().clone().requires_grad_(True)
    f = lambda zz: mlp(model, net_input(model, zz, t, ys), params)
    zdot = f(z)
    B = z.shape[1]
    zero = torch.zeros(1, B, dtype=u.dtype)
    if mode == TEST:
        # D' one-hot pullbacks (utils.jl:35-54); trace = sum_i (e_i^T J)_i
        tr = torch.zeros(B, dtype=u.dtype)
        for i in range(d):
            ct = torch.zeros_like(zdot)
            ct[i, :] = 1.0
            (row,) = torch.autograd.grad(zdot, z, ct, retain_graph=True, create_graph=create_graph)
            tr = tr + row[i, :]
        return torch.cat([zdot, -tr[None, :], zero, zero], dim=0)
    assert eps is not None and eps.shape == (d, B)
    (eJ,) = torch.autograd.grad(zdot, z, eps.to(u.dtype), retain_graph=True, create_graph=create_graph)
    ldot = -(eJ * eps).sum(dim=0, keepdim=True)
    if mode == TRAIN_REG:
        E = _col_norm(zdot, model.reg_squared)[None, :] if model.lam1 != 0 else zero
        n = _col_norm(eJ, model.reg_squared)[None, :] if model.lam2 != 0 else zero
    else:
        E, n = zero, zero
    return torch.cat([zdot, ldot, E, n], dim=0)


# --------------------------------------------------------------------------
# RHS, statement 2: closed forms, no autograd inside

def rhs_closed(model: OracleICNF, mode: int, u: torch.Tensor, theta: torch.Tensor, t,
               eps: Optional[torch.Tensor], ys: Optional[torch.Tensor] = None) -> torch.Tensor:
    d = model.d
    params = unpack_params(theta, model.sizes)
    L = len(params)
    z = u[:d, :]
    B = z.shape[1]
    h = net_input(model, z, t, ys)
    pre = []
    for l, (W, b) in enumerate(params):
        a = W @ h + b[:, None]
        pre.append(a)
        h = act(a, model.activation) if l < L - 1 else a
    zdot = h
    zero = torch.zeros(1, B, dtype=u.dtype)
    if mode == TEST:
        # tr J = tr(W_L D_{L-1} ... D_1 W_1[:, :d]); propagate the d columns
        M = params[0][0][:, :d].unsqueeze(0).expand(B, -1, -1)          # B x n1 x d
        for l in range(1, L):
            D = act_d1(pre[l - 1], model.activation).transpose(0, 1)      # B x n_{l}
            M = params[l][0].unsqueeze(0) @ (D.unsqueeze(2) * M)
        tr = torch.diagonal(M, dim1=1, dim2=2).sum(dim=1)
        return torch.cat([zdot, -tr[None, :], zero, zero], dim=0)
    g = eps.to(u.dtype)
    for l in range(L - 1, -1, -1):
        v = params[l][0].transpose(0, 1) @ g
        if l > 0:
            g = v * act_d1(pre[l - 1], model.activation)
    eJ = v[:d, :]
    ldot = -(eJ * eps).sum(dim=0, keepdim=True)
    if mode == TRAIN_REG:
        E = _col_norm(zdot, model.reg_squared)[None, :] if model.lam1 != 0 else zero
        n = _col_norm(eJ, model.reg_squared)[None, :] if model.lam2 != 0 else zero
    else:
        E, n = zero, zero
    return torch.cat([zdot, ldot, E, n], dim=0)


# --------------------------------------------------------------------------
# Tsit5.  The tableau digits are OrdinaryDiffEq's Tsit5
# [3P, from memory]; tests/test_oracle_tsit5.py checks row sums, order
# conditions and measured convergence order.

C = (0.0, 0.161, 0.327, 0.9, 0.9800255409045097, 1.0, 1.0)
A = (
    (),
    (0.161,),
    (-0.008480655492356989, 0.335480655492357),
    (2.8971530571054935, -6.359448489975075, 4.3622954328695815),
    (5.325864828439257, -11.748883564062828, 7.4955393428898365, -0.09249506636175525),
    (5.86145544294642, -12.92096931784711, 8.159367898576159, -0.071584973281401, -0.028269050394068383),
    (0.09646076681806523, 0.01, 0.4798896504144996, 1.379008574103742, -3.290069515436081, 2.324710524099774),
)
BT = (-0.00178001105222577714, -0.0008164344596567469, 0.007880878010261995,
      -0.1447110071732629, 0.5823571654525552, -0.45808210592918697, 0.015151515151515152)


@dataclass
class SolverOpts:
    """The subset of ``sol_kwargs`` (icnf.jl:84-102) the path uses, for Tsit5."""

    alg: str = "tsit5"              # "tsit5" (north_star) or "vcabm" (the reference's default alg, icnf.jl:89)
    adaptive: bool = True
    dt: float = 0.0                 # fixed step (adaptive=False) or initial dt (0 = automatic)
    reltol: float = 1e-4            # icnf.jl:87
    abstol: float = 1e-4            # icnf.jl:88
    max_steps: int = 100000
    # PI controller constants of OrdinaryDiffEq for Tsit5 [3P, from memory]
    beta1: float = 7.0 / 50.0
    beta2: float = 2.0 / 25.0
    gamma: float = 0.9
    qmin: float = 0.2
    qmax: float = 10.0
    qsteady_min: float = 1.0
    qsteady_max: float = 1.2
    qoldinit: float = 1e-4


@dataclass
class SolveStats:
    naccept: int = 0
    nreject: int = 0
    nf: int = 0
    ts: List[float] = field(default_factory=list)    # accepted step starts
    dts: List[float] = field(default_factory=list)   # accepted step sizes


def _rms(x: torch.Tensor) -> float:
    return float(torch.sqrt((x.detach() ** 2).mean()))


def tsit5_step(f: Callable, u: torch.Tensor, k1: torch.Tensor, t: float, dt: float):
    """One explicit step.  Returns (u_new, k7, err_vector) with
    err = dt * sum_i btilde_i k_i; k7 = f(u_new, t+dt) (FSAL)."""
    ks = [k1]
    for i in range(1, 6):
        ui = u
        for j, a in enumerate(A[i]):
            ui = ui + (dt * a) * ks[j]
        ks.append(f(ui, t + C[i] * dt))
    un = u
    for j, a in enumerate(A[6]):
        un = un + (dt * a) * ks[j]
    k7 = f(un, t + dt)
    ks.append(k7)
    err = None
    for j, bt in enumerate(BT):
        term = (dt * bt) * ks[j]
        err = term if err is None else err + term
    return un, k7, err


def _initial_dt(f, u0, f0, t0, tdir, opts: "SolverOpts", tend, order: int = 5) -> float:
    """Hairer/Wanner starting step as OrdinaryDiffEq uses it [3P, from memory]."""
    sk = opts.abstol + u0.detach().abs() * opts.reltol
    d0 = _rms(u0 / sk)
    d1 = _rms(f0 / sk)
    dt0 = 1e-6 if (d0 < 1e-5 or d1 < 1e-5) else 0.01 * d0 / d1
    dt0 = min(dt0, abs(tend - t0))
    u1 = u0 + (tdir * dt0) * f0
    f1 = f(u1, t0 + tdir * dt0)
    d2 = _rms((f1 - f0) / sk) / dt0
    m = max(d1, d2)
    if m <= 1e-15:
        dt1 = max(1e-6, dt0 * 1e-3)
    else:
        dt1 = 10.0 ** (-(2.0 + math.log10(m)) / (order + 1))
    return min(100.0 * dt0, dt1, abs(tend - t0))


def tsit5_solve(f: Callable, u0: torch.Tensor, t0: float, t1: float, opts: SolverOpts,
                stats: Optional[SolveStats] = None) -> torch.Tensor:
    """Integrate du/dt = f(u, t) from t0 to t1 (t1 < t0 allowed) and return u(t1).

    Fixed step: n = ceil(|t1-t0|/dt) steps, the last one clipped to land on t1.
    Adaptive: error vector scaled by abstol + max(|u|,|u_new|)*reltol, RMS over
    ALL S*B entries (one dt for the whole batch, SURVEY 7 'batch-coupled'), PI
    controller; dt arithmetic is detached from autograd (the accepted step
    sizes are constants of the discretise-then-optimise gradient)."""
    stats = stats if stats is not None else SolveStats()
    tdir = 1.0 if t1 >= t0 else -1.0
    span = abs(t1 - t0)
    u = u0
    t = float(t0)
    if span == 0.0:
        return u
    k1 = f(u, t)
    stats.nf += 1
    if not opts.adaptive:
        assert opts.dt > 0
        n = int(math.ceil(span / opts.dt - 1e-9))
        for i in range(n):
            h = min(opts.dt, span - i * opts.dt)
            h = tdir * h
            u, k1, _ = tsit5_step(f, u, k1, t, h)
            stats.nf += 6
            stats.ts.append(t)
            stats.dts.append(h)
            stats.naccept += 1
            t = t0 + tdir * min(span, (i + 1) * opts.dt)
        return u
    if opts.dt > 0:
        dt = min(opts.dt, span)
    else:
        dt = _initial_dt(f, u, k1, t, tdir, opts, t1)
        stats.nf += 1
    qold = opts.qoldinit
    steps = 0
    while True:
        remaining = abs(t1 - t)
        if remaining <= 1e-12 * max(1.0, abs(t1)):
            break
        last = dt >= remaining * (1.0 - 1e-6)
        h = remaining if last else dt
        un, k7, err = tsit5_step(f, u, k1, t, tdir * h)
        stats.nf += 6
        sk = opts.abstol + torch.maximum(u.detach().abs(), un.detach().abs()) * opts.reltol
        eest = _rms(err / sk)
        steps += 1
        if steps > opts.max_steps:
            raise RuntimeError("tsit5: max_steps exceeded")
        if not math.isfinite(eest):
            raise FloatingPointError("tsit5: non-finite error estimate")
        q11 = eest ** opts.beta1 if eest > 0 else 0.0
        q = q11 / (qold ** opts.beta2)
        q = max(1.0 / opts.qmax, min(1.0 / opts.qmin, q / opts.gamma))
        if eest <= 1.0:
            stats.naccept += 1
            stats.ts.append(t)
            stats.dts.append(tdir * h)
            t = t1 if last else t + tdir * h
            u, k1 = un, k7
            if opts.qsteady_min <= q <= opts.qsteady_max:
                q = 1.0
            qold = max(eest, opts.qoldinit)
            dt = h / q
        else:
            stats.nreject += 1
            dt = h / min(1.0 / opts.qmin, q11 / opts.gamma)
    return u



# --------------------------------------------------------------------------
# VCABM: variable-step, variable-order Adams-Bashforth-Moulton PECE, the reference's default ``alg``
# (icnf.jl:89, third-party OrdinaryDiffEqAdamsBashforthMoulton 2 -- not vendored).  Restated from the published
# algorithm it implements: Hairer, Norsett, Wanner, "Solving ODEs I", III.5 (variable step size Adams formulas in
# terms of the modified divided differences Phi_j, Phi*_j and the coefficients beta_j, g_j) with Shampine-Gordon's
# order selection (error estimates of the neighbouring orders with the constant-step coefficients gamma*_j).
# [3P, from memory, unverifiable here]: which of these quantities OrdinaryDiffEq evaluates at the predicted and
# which at the corrected state, its rule for the first steps (order raised by one per step up to 3) and the
# I-controller constants.  PARITY UNPINNED, like the Tsit5 controller.

GAMMA_STAR = (1.0, -1.0 / 2, -1.0 / 12, -1.0 / 24, -19.0 / 720, -3.0 / 160, -863.0 / 60480, -275.0 / 24192,
              -33953.0 / 3628800, -8183.0 / 1036800, -3250433.0 / 479001600, -4671.0 / 788480, -13695779093.0 / 2615348736000)
VCABM_MAX_ORDER = 12


def vcabm_coefficients(dts: List[float], k: int, kk: int):
    """beta_j (j < kk) and g_j (j <= k) for the step dts[0] after the earlier steps dts[1], dts[2], ...
    HNW III.5: beta_0 = 1, beta_j = beta_{j-1} (t_{n+1} - t_{n-j+1}) / (t_n - t_{n-j});
    c_{0,q} = 1/q, c_{1,q} = 1/(q (q+1)), c_{j,q} = c_{j-1,q} - c_{j-1,q+1} h_n / (t_{n+1} - t_{n-j+1}); g_j = c_{j,1}."""
    h = dts[0]
    beta = [1.0]
    xi, xi0 = h, 0.0
    for j in range(1, kk):
        xi0 += dts[j]
        beta.append(beta[-1] * xi / xi0)
        xi += dts[j]
    c_prev = [1.0 / q for q in range(1, k + 3)]            # c_{0,q}, q = 1 ..
    g = [c_prev[0]]
    if k >= 1:
        c_cur = [1.0 / (q * (q + 1)) for q in range(1, k + 2)]
        g.append(c_cur[0])
        span = h
        for j in range(2, k + 1):
            span += dts[j - 1]                              # t_{n+1} - t_{n-j+1}
            c_next = [c_cur[q] - c_cur[q + 1] * h / span for q in range(len(c_cur) - 1)]
            g.append(c_next[0])
            c_cur = c_next
    return beta, g


def vcabm_solve(f: Callable, u0: torch.Tensor, t0: float, t1: float, opts: SolverOpts,
                stats: Optional[SolveStats] = None) -> torch.Tensor:
    """Adaptive Adams PECE from t0 to t1 (t1 < t0 allowed): u(t1).  Error norm, scaling and acceptance as in
    tsit5_solve (whole-batch RMS); step size by the I-controller q = EEst^(1/(k+1)) / gamma."""
    stats = stats if stats is not None else SolveStats()
    tdir = 1.0 if t1 >= t0 else -1.0
    span = abs(t1 - t0)
    u, t = u0, float(t0)
    if span == 0.0:
        return u
    fn = f(u, t)
    stats.nf += 1
    if opts.dt > 0:
        dt = min(opts.dt, span)
    else:
        dt = _initial_dt(f, u, fn, t, tdir, opts, t1, order=1)
        stats.nf += 1
    k, nstep = 1, 0
    hist = [0.0] * (VCABM_MAX_ORDER + 2)        # hist[i] = i-th previous accepted (signed) step
    phistar_prev: List[torch.Tensor] = []        # Phi*_j(n-1)
    attempts = 0

    def nrm(x, sk):
        return _rms(x / sk)

    while True:
        remaining = abs(t1 - t)
        if remaining <= 1e-12 * max(1.0, abs(t1)):
            break
        last = dt >= remaining * (1.0 - 1e-6)
        h = tdir * (remaining if last else dt)
        attempts += 1
        if attempts > opts.max_steps:
            raise RuntimeError("vcabm: max_steps exceeded")
        kk = min(k + 1, nstep + 1)
        dts = [h] + hist
        beta, g = vcabm_coefficients(dts, k, kk)
        phi = [fn]
        phistar = [fn]
        for j in range(1, kk):
            phi.append(phi[j - 1] - phistar_prev[j - 1])
            phistar.append(beta[j] * phi[j])
        p = u
        for j in range(k):
            p = p + (h * g[j]) * phistar[j]
        fp = f(p, t + h)
        stats.nf += 1
        php = [fp]
        for j in range(1, kk + 1):
            php.append(php[j - 1] - phistar[j - 1])
        un = p + (h * g[k]) * php[k]
        sk = opts.abstol + torch.maximum(u.detach().abs(), un.detach().abs()) * opts.reltol
        eest = nrm((h * (g[k] - g[k - 1])) * php[k], sk)
        if not math.isfinite(eest):
            raise FloatingPointError("vcabm: non-finite error estimate")
        if eest > 1.0:
            stats.nreject += 1
            q = eest ** (1.0 / (k + 1)) / opts.gamma
            dt = abs(h) / min(1.0 / opts.qmin, max(1.0 / opts.qmax, q))
            continue
        # accepted: final evaluation (PECE), order for the next step
        fnew = f(un, t1 if last else t + h)
        stats.nf += 1
        knew = k
        if nstep + 1 <= 4 or k < 3:
            knew = min(k + 1, 3, VCABM_MAX_ORDER)
        else:
            errm1 = nrm((h * GAMMA_STAR[k - 1]) * php[k - 1], sk)
            errm2 = nrm((h * GAMMA_STAR[k - 2]) * php[k - 2], sk)
            if max(errm1, errm2) <= eest:
                knew = k - 1
            elif k < VCABM_MAX_ORDER and kk >= k + 1:
                errp1 = nrm((h * GAMMA_STAR[k + 1]) * php[k + 1], sk)
                if errp1 < eest:
                    knew = k + 1
                    eest = errp1
        stats.naccept += 1
        stats.ts.append(t)
        stats.dts.append(h)
        t = t1 if last else t + h
        u, fn = un, fnew
        phistar_prev = phistar
        hist = [h] + hist[:-1]
        nstep += 1
        k = knew
        q = eest ** (1.0 / (k + 1)) / opts.gamma if eest > 0 else 0.0
        q = max(1.0 / opts.qmax, min(1.0 / opts.qmin, q))
        if opts.qsteady_min <= q <= opts.qsteady_max:
            q = 1.0
        dt = abs(h) / q
    return u

# --------------------------------------------------------------------------
# problem build / readout / API (base_icnf.jl)

LOG_2PI = math.log(2.0 * math.pi)


def make_u0(model: OracleICNF, xs: torch.Tensor) -> torch.Tensor:
    """u0 = [xs; zeros(naug + 3, B)] (base_icnf.jl:254-265)."""
    B = xs.shape[1]
    return torch.cat([xs, torch.zeros(model.naug + 3, B, dtype=xs.dtype)], dim=0)


def _rhs(model, mode, theta, eps, ys, closed, create_graph):
    if closed:
        return lambda u, t: rhs_closed(model, mode, u, theta, t, eps, ys)
    return lambda u, t: rhs_ad(model, mode, u, theta, t, eps, ys, create_graph=create_graph)


def solve(model: OracleICNF, mode: int, u0, theta, eps, ys=None, t0=None, t1=None,
          opts: Optional[SolverOpts] = None, closed: bool = True,
          create_graph: bool = False, stats: Optional[SolveStats] = None) -> torch.Tensor:
    """``base_sol`` (base_icnf.jl:134-140): u(t1), final state only."""
    opts = opts or SolverOpts()
    t0 = model.tspan[0] if t0 is None else t0
    t1 = model.tspan[1] if t1 is None else t1
    solver = vcabm_solve if opts.alg == "vcabm" else tsit5_solve
    return solver(_rhs(model, mode, theta, eps, ys, closed, create_graph), u0, t0, t1, opts, stats)


def readout(model: OracleICNF, mode: int, fsol: torch.Tensor):
    """``inference_sol`` (base_icnf.jl:158-172) + ``reg_z_aug`` (:106-132)."""
    d = model.d
    z = fsol[:d, :]
    dlogp = fsol[d, :]
    E = fsol[d + 1, :]
    n = fsol[d + 2, :]
    logpz = -0.5 * d * LOG_2PI - 0.5 * (z * z).sum(dim=0)
    logpx = logpz - dlogp
    if mode == TRAIN_REG and model.naug > 0 and model.lam3 != 0:
        Adot = _col_norm(z[model.nvars:, :], model.reg_squared)
    else:
        Adot = torch.zeros_like(dlogp)
    return logpx, (E, n, Adot)


def inference(model: OracleICNF, mode: int, xs, theta, eps, ys=None, t1=None,
              opts: Optional[SolverOpts] = None, closed: bool = True,
              create_graph: bool = False, stats: Optional[SolveStats] = None):
    """``inference`` (base_icnf.jl:406-424).  ``t1`` overrides tspan[2] (STEER)."""
    fsol = solve(model, mode, make_u0(model, xs), theta, eps, ys, None, t1, opts, closed, create_graph, stats)
    return readout(model, mode, fsol)


def generate(model: OracleICNF, mode: int, z0, theta, eps, ys=None, t1=None,
             opts: Optional[SolverOpts] = None, closed: bool = True,
             stats: Optional[SolveStats] = None) -> torch.Tensor:
    """``generate`` (base_icnf.jl:351-404, :185-194): z0 ~ basedist is supplied
    by the caller (D' x n); integrates t1 -> t0; returns the first nvars rows."""
    n = z0.shape[1]
    u0 = torch.cat([z0, torch.zeros(3, n, dtype=z0.dtype)], dim=0)
    t_hi = model.tspan[1] if t1 is None else t1
    fsol = solve(model, mode, u0, theta, eps, ys, t_hi, model.tspan[0], opts, closed, False, stats)
    return fsol[: model.nvars, :]


def loss(model: OracleICNF, mode: int, xs, theta, eps, ys=None, t1=None,
         opts: Optional[SolverOpts] = None, closed: bool = True,
         create_graph: bool = False, stats: Optional[SolveStats] = None) -> torch.Tensor:
    """``loss`` for ICNF in matrix mode (icnf.jl:628-649)."""
    logpx, (E, n, Adot) = inference(model, mode, xs, theta, eps, ys, t1, opts, closed, create_graph, stats)
    return (-logpx + model.lam1 * E + model.lam2 * n + model.lam3 * Adot).mean()


def loss_grad(model: OracleICNF, mode: int, xs, theta, eps, ys=None, t1=None,
              opts: Optional[SolverOpts] = None, want_dxs: bool = False,
              stats: Optional[SolveStats] = None):
    """Loss and its gradient w.r.t. theta (and xs) by reverse-mode AD THROUGH
    the discrete solve (discretise-then-optimise).  The reference uses a
    continuous QuadratureAdjoint at tol 1e-4 instead (icnf.jl:90-99) [3P]; the
    two agree to solver tolerance, not to rounding (hard parts)."""
    theta = theta.detach().clone().requires_grad_(True)
    xs = xs.detach().clone().requires_grad_(want_dxs)
    val = loss(model, mode, xs, theta, eps, ys, t1, opts, closed=False, create_graph=True, stats=stats)
    grads = torch.autograd.grad(val, [theta] + ([xs] if want_dxs else []))
    return val.detach(), grads[0], (grads[1] if want_dxs else None)


def steer_t1(model: OracleICNF, mode: int, r: float) -> float:
    """``steer_tspan`` (base_icnf.jl:23-43): t1 + |t1-t0| r, r ~ U(-rate, rate),
    TrainMode{true} and steer_rate != 0 only; ``r`` is supplied by the caller."""
    t0, t1 = model.tspan
    if mode == TRAIN_REG and model.steer_rate != 0:
        return t1 + abs(t1 - t0) * r
    return t1


# --------------------------------------------------------------------------
# continuous adjoint: what the reference's sensealg computes

def loss_grad_continuous(model: OracleICNF, mode: int, xs, theta, eps, ys=None, t1=None,
                         opts: Optional[SolverOpts] = None, adj_opts: Optional[SolverOpts] = None,
                         want_dxs: bool = False, stats: Optional[SolveStats] = None):
    """Loss and its gradient by the CONTINUOUS adjoint -- the semantics of the
    reference's ``sensealg = QuadratureAdjoint(autojacvec = ZygoteVJP())``
    (/root/reference/src/core/icnf.jl:90-99) [3P]: solve u forward, then integrate

        d lam / dt = -(df/du)' lam,        lam(t1) = dL/du(t1)
        d mu  / dt = -(df/dtheta)' lam,    mu(t1)  = 0

    from t1 back to t0; dL/dtheta = mu(t0), dL/dxs = lam(t0)[1:nvars].  The
    vector-Jacobian products go THROUGH ``rhs_ad`` (second-order reverse mode, as
    Zygote-over-Lux.vector_jacobian_product does in the reference).

    QuadratureAdjoint interpolates the stored forward solution and evaluates the
    theta integral by Gauss-Kronrod quadrature; here u is re-integrated backwards
    together with (lam, mu) in one Tsit5 solve (``adj_opts``, default = ``opts``).
    Both discretise the same continuous equations, so they agree to their solver
    tolerances; with tight tolerances this function returns the exact gradient of
    the exact flow, which is what tests/test_oracle_adjoint.py uses it for: it
    measures how far the product's discretise-then-optimise gradient is from it.
    """
    opts = opts or SolverOpts()
    adj_opts = adj_opts or opts
    theta = theta.detach()
    t0 = model.tspan[0]
    t_end = model.tspan[1] if t1 is None else t1
    with torch.no_grad():
        u1 = solve(model, mode, make_u0(model, xs.detach()), theta, eps, ys, t0, t_end, opts, closed=True, stats=stats)
    u1 = u1.detach().clone().requires_grad_(True)
    logpx, (E, n, Adot) = readout(model, mode, u1)
    val = (-logpx + model.lam1 * E + model.lam2 * n + model.lam3 * Adot).mean()
    (lam1_,) = torch.autograd.grad(val, u1)
    S, B = u1.shape
    P = theta.numel()

    def pack(u, lam, mu):
        return torch.cat([u.reshape(-1), lam.reshape(-1), mu.reshape(-1)])

    def g(y, t):
        u = y[: S * B].reshape(S, B).detach().clone().requires_grad_(True)
        lam = y[S * B: 2 * S * B].reshape(S, B).detach()
        th = theta.clone().requires_grad_(True)
        with torch.enable_grad():
            fu = rhs_ad(model, mode, u, th, t, eps, ys, create_graph=True)
            gu, gth = torch.autograd.grad(fu, [u, th], lam, allow_unused=True)
        gu = torch.zeros_like(u) if gu is None else gu
        gth = torch.zeros_like(th) if gth is None else gth
        return pack(fu.detach(), -gu.detach(), -gth.detach())

    y1 = pack(u1.detach(), lam1_.detach(), torch.zeros(P, dtype=theta.dtype))
    adj_stats = SolveStats()
    y0 = tsit5_solve(g, y1, t_end, t0, adj_opts, adj_stats)
    lam0 = y0[S * B: 2 * S * B].reshape(S, B)
    mu0 = y0[2 * S * B:]
    if stats is not None:
        stats.nf += adj_stats.nf
    return val.detach(), mu0, (lam0[: model.nvars, :] if want_dxs else None)
