"""Host-side mirror of the reference's flow API for the B200 path.

Same names, argument order and meaning as ContinuousNormalizingFlows.jl
(/root/reference/src/core/base_icnf.jl:406-523, src/core/icnf.jl:53-141,
:628-649): ``ICNF(...)``, ``inference``, ``generate``, ``loss``, the callable
layer, ``TestMode`` / ``TrainMode``.  Matrices keep the reference's logical
shape ``R x B`` (one column per sample).  Every numeric result comes from the
CUDA library through the C ABI (``_lib``); nothing is computed here.

Accepted array types:
  * numpy arrays  -> host entry points (``icnf_inference`` ...): the library
    copies in and out, as it would for Julia ``Array``s passed by ``ccall``;
  * torch CUDA tensors -> ``_dev`` entry points on torch's current stream, no
    copies, results returned as torch tensors.

Python identifiers cannot carry the reference's subscripted lambdas; they are
spelled ``lambda1``, ``lambda2``, ``lambda3`` here.
"""

from __future__ import annotations

import ctypes as C
import math
from dataclasses import dataclass, field
from typing import Any, Optional, Sequence, Tuple

import numpy as np

from . import _lib
from ._lib import ICNFError, lib

try:  # torch is plumbing only (device memory, streams); the host path works without it
    import torch
except Exception:  # pragma: no cover
    torch = None


# ---------------------------------------------------------------- modes (src/core/types.jl)

def _alg_code(alg):
    """``sol_kwargs.alg`` (icnf.jl:89) -> icnf_alg: 0 = Tsit5, 1 = VCABM, None = not served."""
    name = str(alg).lower().rstrip("()")
    return {"tsit5": 0, "vcabm": 1}.get(name)


class Mode:
    code: int = -1


class TestMode(Mode):
    __test__ = False  # not a pytest class
    code = _lib.MODE_TEST

    def __repr__(self):
        return "TestMode()"


class TrainMode(Mode):
    """``TrainMode{REG}``; ``TrainMode()`` is ``TrainMode{true}`` (types.jl:5-7)."""

    def __init__(self, reg: bool = True):
        self.reg = bool(reg)
        self.code = _lib.MODE_TRAIN_REG if self.reg else _lib.MODE_TRAIN_NOREG

    def __repr__(self):
        return f"TrainMode({self.reg})"


class ComputeMode:
    pass


class MatrixMode(ComputeMode):
    pass


class B200MatrixMode(MatrixMode):
    """The compute mode this package adds: batched (matrix) evaluation on a B200
    through libicnf_b200.so.  It takes the place of ``LuxVecJacMatrixMode`` /
    ``DIVecJacMatrixMode`` (types.jl:22-35) behind the ``compute_mode`` switch."""

    def __repr__(self):
        return "B200MatrixMode()"


# ---------------------------------------------------------------- network description (Lux.Chain of Dense)
@dataclass(frozen=True)
class Dense:
    n_in: int
    n_out: int
    activation: str = "identity"


@dataclass(frozen=True)
class PlanarLayer:
    """``PlanarLayer(n_in => n_out, activation; use_bias)`` (src/layers/planar_layer.jl:1-97): f(x) = u * act(w'x + b),
    parameters ``(u [n_out], w [n_in], b [1])`` in that order.  It IS ``Dense(n_in => 1, act)`` followed by a bias-free
    ``Dense(1 => n_out)``, which is how it reaches the kernels: the mirror (and the Julia glue) re-orders the parameters
    into ``[w; b; u; 0]`` on the way in and the gradient back on the way out; the output bias stays zero."""
    n_in: int
    n_out: int
    activation: str = "identity"
    use_bias: bool = True

    @property
    def n_params(self) -> int:
        return self.n_out + self.n_in + int(self.use_bias)

    def to_lib(self, ps):
        """planar (u, w[, b]) -> [W1 = w; b1 = b; W2 = u; b2 = 0] (numpy or torch, any device)."""
        no, ni = self.n_out, self.n_in
        if _is_torch(ps):
            p = ps.detach().to(dtype=torch.float32).reshape(-1)
            b = p[no + ni:no + ni + 1] if self.use_bias else p.new_zeros(1)
            return torch.cat([p[no:no + ni], b, p[:no], p.new_zeros(no)])
        p = np.asarray(ps, dtype=np.float32).reshape(-1)
        b = p[no + ni:no + ni + 1] if self.use_bias else np.zeros(1, np.float32)
        return np.concatenate([p[no:no + ni], b, p[:no], np.zeros(no, np.float32)])

    def grad_from_lib(self, g):
        """gradient w.r.t. [W1; b1; W2; b2] -> gradient w.r.t. (u, w[, b])."""
        no, ni = self.n_out, self.n_in
        parts = [g[ni + 1:ni + 1 + no], g[:ni]] + ([g[ni:ni + 1]] if self.use_bias else [])
        return torch.cat(parts) if _is_torch(g) else np.concatenate(parts)


@dataclass(frozen=True)
class Chain:
    layers: Tuple[Dense, ...]

    def __init__(self, *layers: Dense):
        object.__setattr__(self, "layers", tuple(layers))


@dataclass
class SolverStats:
    naccept: int = 0
    nreject: int = 0
    nf: int = 0
    status: int = 0
    t_final: float = 0.0
    dt_last: float = 0.0


DEFAULT_SOL_KWARGS = dict(
    alg="Tsit5",          # the stepper this path implements (BASELINE.json north_star)
    adaptive=True,
    dt=0.0,               # fixed step when adaptive=False; initial step (0 = automatic) otherwise
    reltol=1e-4,          # icnf.jl:87
    abstol=1e-4,          # icnf.jl:88
    maxiters=100000,      # the reference uses typemax(Int) (icnf.jl:86)
    save_everystep=False,
)


def _is_torch(x) -> bool:
    return torch is not None and isinstance(x, torch.Tensor)


class ICNF:
    """Mirror of ``ICNF(; ...)`` (src/core/icnf.jl:53-141) bound to one GPU."""

    def __init__(self, *, data_type=np.float32, compute_mode: ComputeMode = None, inplace: bool = False,
                 autonomous: bool = False, device: int = 0, rng: Any = None, tspan: Tuple[float, float] = (0.0, 1.0),
                 nvariables: int = 1, naugments: Optional[int] = None, nconditions: int = 0,
                 n_in: Optional[int] = None, n_out: Optional[int] = None, n_hidden: Optional[int] = None,
                 nn: Optional[Chain] = None, steer_rate: float = 0.1, lambda1: float = 0.01, lambda2: float = 0.01,
                 lambda3: float = 0.01, epsdist: str = "gaussian", reg_squared: bool = False,
                 precision: str = "fp32", sol_kwargs: Optional[dict] = None):
        if np.dtype(data_type) != np.float32:
            raise ValueError("the B200 path computes in Float32 (the reference's default data_type, icnf.jl:54)")
        self.data_type = np.float32
        self.compute_mode = compute_mode if compute_mode is not None else B200MatrixMode()
        if not isinstance(self.compute_mode, B200MatrixMode):
            raise ValueError("this package implements B200MatrixMode only")
        self.inplace = bool(inplace)          # kept for signature parity; results are identical either way
        self.autonomous = bool(autonomous)
        self.device = int(device)
        self.rng = rng if isinstance(rng, np.random.Generator) else np.random.default_rng(rng)
        self.tspan = (float(tspan[0]), float(tspan[1]))
        self.nvariables = int(nvariables)
        self.naugments = int(nvariables + 1 if naugments is None else naugments)       # icnf.jl:62
        self.nconditions = int(nconditions)
        d = self.nvariables + self.naugments
        n_in = d + (0 if self.autonomous else 1) + self.nconditions if n_in is None else int(n_in)   # icnf.jl:64
        n_out = d if n_out is None else int(n_out)
        n_hidden = 4 * n_in if n_hidden is None else int(n_hidden)                     # icnf.jl:66
        if nn is None:                                                                  # icnf.jl:67-71
            nn = Chain(Dense(n_in, n_hidden, "softplus"), Dense(n_hidden, n_hidden, "softplus"), Dense(n_hidden, n_out))
        self.nn = nn
        self.steer_rate = float(steer_rate)
        self.lambda1, self.lambda2, self.lambda3 = float(lambda1), float(lambda2), float(lambda3)
        if epsdist not in ("gaussian", "rademacher"):
            raise ValueError("epsdist must be 'gaussian' (reference default, icnf.jl:80-83) or 'rademacher'")
        self.epsdist = epsdist
        self.sol_kwargs = dict(DEFAULT_SOL_KWARGS)
        if sol_kwargs:
            self.sol_kwargs.update(sol_kwargs)
        if _alg_code(self.sol_kwargs.get("alg", "Tsit5")) is None:
            raise ValueError("the B200 path integrates with Tsit5 (north_star) or VCABM (the reference's default alg, served by "
                             "the single-launch solves for inference / generate / loss); pass alg='Tsit5' or alg='VCABM'")

        self.planar = None
        if isinstance(nn, PlanarLayer):
            nn = Chain(nn)
        if isinstance(nn, Chain) and len(nn.layers) == 1 and isinstance(nn.layers[0], PlanarLayer):
            self.planar = nn.layers[0]
            if (self.planar.n_in, self.planar.n_out) != (n_in, n_out):
                raise ValueError(f"network must map {n_in} -> {n_out}")
            self.nn = nn
            nn = Chain(Dense(n_in, 1, self.planar.activation), Dense(1, n_out))
        sizes, acts = self._check_chain(nn, n_in, n_out)
        self.sizes = sizes
        cfg = _lib.Config()
        cfg.abi_version = _lib.ICNF_ABI_VERSION
        cfg.nvars, cfg.naug, cfg.ncond = self.nvariables, self.naugments, self.nconditions
        cfg.autonomous = int(self.autonomous)
        cfg.n_layers = len(sizes) - 1
        for i, s in enumerate(sizes):
            cfg.sizes[i] = s
        cfg.activation = _lib.ACT[acts]
        cfg.lambda1, cfg.lambda2, cfg.lambda3 = self.lambda1, self.lambda2, self.lambda3
        cfg.reg_squared = int(bool(reg_squared))
        cfg.precision = _lib.PRECISION[precision]
        cfg.device = self.device
        self._cfg = cfg
        self._h = C.c_void_p()
        rc = lib.icnf_create(C.byref(cfg), C.byref(self._h))
        if rc != _lib.OK:
            raise ICNFError(rc, lib.icnf_last_error(None).decode())
        self._params_key = None
        self._last_stats = SolverStats()
        self._pending_stats = None    # device-resident icnf_stats of the last asynchronous (_dev) loss/gradient call

    # -- helpers
    @staticmethod
    def _check_chain(nn: Chain, n_in: int, n_out: int):
        if not isinstance(nn, Chain) or not nn.layers:
            raise ValueError("nn must be a Chain of Dense layers")
        if len(nn.layers) > _lib.ICNF_MAX_LAYERS:
            raise ValueError(f"at most {_lib.ICNF_MAX_LAYERS} Dense layers")
        sizes = [nn.layers[0].n_in]
        hidden_acts = set()
        for i, layer in enumerate(nn.layers):
            if layer.n_in != sizes[-1]:
                raise ValueError("Dense sizes do not chain")
            sizes.append(layer.n_out)
            if i < len(nn.layers) - 1:
                hidden_acts.add(layer.activation)
            elif layer.activation != "identity":
                raise ValueError("the last Dense layer must be linear (icnf.jl:70)")
        if sizes[0] != n_in or sizes[-1] != n_out:
            raise ValueError(f"network must map {n_in} -> {n_out}")
        if len(hidden_acts) > 1:
            raise ValueError("all hidden layers must share one activation")
        act = hidden_acts.pop() if hidden_acts else "identity"
        if act not in _lib.ACT:
            raise ValueError(f"unsupported activation {act!r}")
        return tuple(sizes), act

    def __del__(self):
        h = getattr(self, "_h", None)
        if h is not None and h.value and lib is not None:     # `lib` is None during interpreter shutdown
            try:
                lib.icnf_destroy(h)
            except Exception:
                pass
            self._h = C.c_void_p()

    # -- solver statistics.  Host-pointer calls return them synchronously.  The device-pointer loss/gradient call is
    # asynchronous: its icnf_stats record stays on the device until somebody asks (reading it synchronises the
    # stream).  `last_stats` reads it lazily; `check_last()` raises if the device loop failed (max_steps, dt
    # underflow, non-finite state) -- the backward kernels return a NaN gradient in that case, never a stale one.
    @property
    def last_stats(self) -> "SolverStats":
        if self._pending_stats is not None:
            ds, self._pending_stats = self._pending_stats, None
            raw = ds.cpu().numpy()
            f = raw.view(np.float32)
            self._last_stats = SolverStats(int(raw[0]), int(raw[1]), int(raw[2]), int(raw[3]), float(f[4]), float(f[5]))
        return self._last_stats

    @last_stats.setter
    def last_stats(self, st: "SolverStats"):
        self._pending_stats = None
        self._last_stats = st

    def check_last(self) -> "SolverStats":
        """Statistics of the most recent solve; raises ``ICNFError`` if its device loop did not reach t1."""
        st = self.last_stats
        if st.status != _lib.OK:
            raise ICNFError(st.status, f"tsit5 device loop failed (t = {st.t_final}, accepted {st.naccept}, rejected {st.nreject})")
        return st

    @property
    def kernel_family(self) -> str:
        return lib.icnf_kernel_family(self._h).decode()

    def solve_path(self, mode) -> str:
        """where a solve of ``mode`` runs: 'tiny', 'narrow' (single launch, any narrow shape), 'generic' or 'tc'"""
        return lib.icnf_solve_path(self._h, mode.code).decode()

    @property
    def n_params(self) -> int:
        """length of ``ps`` as the caller sees it (a PlanarLayer has fewer entries than the Dense pair that serves it)"""
        return self.planar.n_params if self.planar is not None else int(lib.icnf_n_params(self._h))

    @property
    def n_params_lib(self) -> int:
        return int(lib.icnf_n_params(self._h))

    @property
    def launch_count(self) -> int:
        return int(lib.icnf_launch_count(self._h))

    def set_profiling(self, enabled: bool):
        self._check(lib.icnf_set_profiling(self._h, int(bool(enabled))))

    def kernel_times_ms(self):
        """device time of the last {forward, loss sum, backward, grad reduce} kernels (profiling on)"""
        buf = (C.c_float * 4)()
        self._check(lib.icnf_kernel_times(self._h, buf))
        return dict(zip(("forward", "loss_sum", "backward", "grad_reduce"), (float(x) for x in buf)))

    def _check(self, rc: int):
        if rc != _lib.OK:
            raise ICNFError(rc, lib.icnf_last_error(self._h).decode())

    def _solver(self, overrides: Optional[dict] = None) -> _lib.Solver:
        kw = dict(self.sol_kwargs)
        if overrides:
            kw.update(overrides)
        s = _lib.Solver()
        s.adaptive = int(bool(kw.get("adaptive", True)))
        s.dt = float(kw.get("dt", 0.0) or 0.0)
        s.reltol = float(kw.get("reltol", 1e-4))
        s.abstol = float(kw.get("abstol", 1e-4))
        mi = kw.get("maxiters", 100000)
        s.max_steps = int(min(mi, 2 ** 31 - 1))
        for k in ("beta1", "beta2", "gamma", "qmin", "qmax", "qsteady_min", "qsteady_max", "qoldinit"):
            setattr(s, k, float(kw.get(k, 0.0)))
        code = _alg_code(kw.get("alg", "Tsit5"))
        if code is None:
            raise ValueError(f"unknown alg {kw.get('alg')!r}: 'Tsit5' or 'VCABM'")
        s.alg = code
        return s

    def _set_params(self, ps):
        """``ps`` travels with every reference call; upload only when it changed."""
        if self.planar is not None:
            ps = self.planar.to_lib(ps)
        if _is_torch(ps):
            p = ps.detach().to(dtype=torch.float32).contiguous()
            if not p.is_cuda:
                p = p.cpu().numpy()
            else:
                # device-resident parameters can be changed behind torch's back (icnf_adam_step_dev
                # updates them in place), so they are handed to the library on every call
                self._check(lib.icnf_set_params_dev(self._h, p.data_ptr(), p.numel(),
                                                    torch.cuda.current_stream(p.device).cuda_stream))
                self._params_key = None
                return
            ps = p
        p = np.ascontiguousarray(np.asarray(ps, dtype=np.float32).reshape(-1))
        key = ("n", p.tobytes())
        if key != self._params_key:
            self._check(lib.icnf_set_params(self._h, p.ctypes.data, p.size))
            self._params_key = key

    def _noise(self, mode: Mode, eps, seed, sample_offset: int):
        """Hutchinson probe source: a supplied matrix, or an in-kernel Philox draw
        seeded from ``icnf.rng`` once per solve (base_icnf.jl:258-259)."""
        n = _lib.Noise()
        n.sample_offset = int(sample_offset)
        if eps is not None or isinstance(mode, TestMode):
            n.kind = _lib.EPS_SUPPLIED
            n.seed = 0
        else:
            n.kind = _lib.EPS[self.epsdist]
            n.seed = int(self.rng.integers(0, 2 ** 63)) if seed is None else int(seed)
        return n

    def steer_tspan(self, mode: Mode, seed: Optional[int] = None) -> Tuple[float, float]:
        """``steer_tspan`` (base_icnf.jl:23-43).  The draw is the library's (``icnf_steer_tspan``: Philox stream
        "STER" keyed on a seed taken from ``icnf.rng``), so identically seeded ranks steer to the same t1."""
        t0, t1 = self.tspan
        if isinstance(mode, TrainMode) and mode.reg and self.steer_rate != 0.0:
            s = int(self.rng.integers(0, 2 ** 63)) if seed is None else int(seed)
            out = C.c_float()
            self._check(lib.icnf_steer_tspan(mode.code, t0, t1, self.steer_rate, s, C.byref(out)))
            t1 = float(out.value)
        return t0, t1

    # callable layer (base_icnf.jl:509-523)
    def __call__(self, xs, ps, st):
        if isinstance(xs, tuple):
            x, y = xs
            return inference(self, TrainMode(False), x, y, ps, st)[0], st
        return inference(self, TrainMode(False), xs, ps, st)[0], st


def measure_fp32_peak(device: int = 0) -> float:
    """FP32 FMA throughput of the device in TFLOP/s (FFMA-chain microbenchmark in the library)."""
    out = C.c_float()
    rc = lib.icnf_measure_fp32_peak(int(device), C.byref(out))
    if rc != _lib.OK:
        raise ICNFError(rc, "fp32 peak measurement failed")
    return float(out.value)


# ---------------------------------------------------------------- parameters
def setup(rng, icnf: ICNF):
    """``LuxCore.setup(rng, icnf)`` + ``ComponentArray(ps)``: a flat Float32 vector
    [vec(W1); b1; vec(W2); b2; ...] (W column-major), glorot-uniform weights and
    zero biases (Lux's Dense defaults), and an empty state."""
    g = rng if isinstance(rng, np.random.Generator) else np.random.default_rng(rng)
    if icnf.planar is not None:      # planar_layer.jl:36-51: u, w ~ init_weight (glorot-uniform), b = 0
        pl = icnf.planar
        lu, lw = math.sqrt(6.0 / (pl.n_out + 1)), math.sqrt(6.0 / (pl.n_in + 1))
        parts = [g.uniform(-lu, lu, size=pl.n_out), g.uniform(-lw, lw, size=pl.n_in)] + ([np.zeros(1)] if pl.use_bias else [])
        return np.concatenate(parts).astype(np.float32), {}
    parts = []
    for layer in icnf.nn.layers:
        lim = math.sqrt(6.0 / (layer.n_in + layer.n_out))
        parts.append(g.uniform(-lim, lim, size=layer.n_in * layer.n_out))
        parts.append(np.zeros(layer.n_out))
    return np.concatenate(parts).astype(np.float32), {}


# ---------------------------------------------------------------- marshalling
class _Arg:
    """One R x B matrix argument as a raw address + keep-alive reference."""

    def __init__(self, x, rows: int, name: str, on_device: bool):
        self.keep = None
        self.ptr = None
        self.B = None
        if x is None:
            return
        if on_device:
            if not _is_torch(x) or not x.is_cuda:
                raise TypeError(f"{name}: mixing host and device arrays in one call is not supported")
            if x.dim() == 1:
                x = x.reshape(rows, -1) if rows == 1 else x.reshape(rows, 1)
            if x.shape[0] != rows:
                raise ValueError(f"{name} must have {rows} rows, got {tuple(x.shape)}")
            t = x.detach().to(torch.float32).t().contiguous()     # (B, R): B records of R floats
            self.keep, self.ptr, self.B = t, t.data_ptr(), t.shape[0]
        else:
            if _is_torch(x):
                x = x.detach().cpu().numpy()
            a = np.asarray(x, dtype=np.float32)
            if a.ndim == 1:
                a = a.reshape(rows, -1) if rows == 1 else a.reshape(rows, 1)
            if a.shape[0] != rows:
                raise ValueError(f"{name} must have {rows} rows, got {a.shape}")
            a = np.asfortranarray(a)
            self.keep, self.ptr, self.B = a, a.ctypes.data, a.shape[1]


def _on_device(*xs) -> bool:
    return any(_is_torch(x) and x.is_cuda for x in xs if x is not None)


def _stream(icnf: ICNF):
    return torch.cuda.current_stream(icnf.device).cuda_stream


def _stats_from(s: _lib.Stats) -> SolverStats:
    return SolverStats(s.naccept, s.nreject, s.nf, s.status, s.t_final, s.dt_last)


def _dev_stats(icnf: ICNF):
    return torch.zeros(6, dtype=torch.int32, device=f"cuda:{icnf.device}")


def _read_dev_stats(icnf: ICNF, t) -> SolverStats:
    raw = t.cpu().numpy()
    f = raw.view(np.float32)
    st = SolverStats(int(raw[0]), int(raw[1]), int(raw[2]), int(raw[3]), float(f[4]), float(f[5]))
    if st.status != _lib.OK:
        raise ICNFError(st.status, f"tsit5 device loop failed (t = {st.t_final}, accepted {st.naccept})")
    return st


def _split(args, n_tail: int, what: str):
    """reference signatures put the optional ``ys`` before the trailing arguments"""
    if len(args) == n_tail:
        return (None,) + tuple(args)
    if len(args) == n_tail + 1:
        return tuple(args)
    raise TypeError(f"{what}: wrong number of arguments")


# ---------------------------------------------------------------- S1 / S2 seams
def augmented_f(icnf: ICNF, mode: Mode, u, ps, t: float, eps=None, ys=None):
    """One RHS evaluation (``augmented_f``, icnf.jl:297-339 / :517-559): u -> du, both S x B."""
    icnf._set_params(ps)
    dev = _on_device(u, eps, ys)
    S = icnf.nvariables + icnf.naugments + 3
    ua = _Arg(u, S, "u", dev)
    ea = _Arg(eps, S - 3, "eps", dev)
    ya = _Arg(ys, icnf.nconditions, "ys", dev)
    if dev:
        out = torch.empty((ua.B, S), dtype=torch.float32, device=ua.keep.device)
        icnf._check(lib.icnf_rhs_dev(icnf._h, mode.code, float(t), ua.ptr, ea.ptr, ya.ptr, out.data_ptr(), ua.B, _stream(icnf)))
        return out.t()
    out = np.empty((S, ua.B), dtype=np.float32, order="F")
    icnf._check(lib.icnf_rhs(icnf._h, mode.code, float(t), ua.ptr, ea.ptr, ya.ptr, out.ctypes.data, ua.B))
    return out


def base_sol(icnf: ICNF, mode: Mode, u0, ps, tspan=None, eps=None, ys=None, seed=None, sample_offset: int = 0, **sol):
    """``base_sol`` (base_icnf.jl:134-140): u(t_end), S x B."""
    icnf._set_params(ps)
    dev = _on_device(u0, eps, ys)
    S = icnf.nvariables + icnf.naugments + 3
    t0, t1 = tspan if tspan is not None else icnf.tspan
    ua = _Arg(u0, S, "u0", dev)
    ea = _Arg(eps, S - 3, "eps", dev)
    ya = _Arg(ys, icnf.nconditions, "ys", dev)
    noise = icnf._noise(mode, eps, seed, sample_offset)
    solver = icnf._solver(sol)
    if dev:
        out = torch.empty((ua.B, S), dtype=torch.float32, device=ua.keep.device)
        ds = _dev_stats(icnf)
        icnf._check(lib.icnf_solve_dev(icnf._h, mode.code, C.byref(solver), t0, t1, ua.ptr, C.byref(noise), ea.ptr, ya.ptr,
                                       out.data_ptr(), ds.data_ptr(), ua.B, _stream(icnf)))
        icnf.last_stats = _read_dev_stats(icnf, ds)
        return out.t()
    out = np.empty((S, ua.B), dtype=np.float32, order="F")
    st = _lib.Stats()
    rc = lib.icnf_solve(icnf._h, mode.code, C.byref(solver), t0, t1, ua.ptr, C.byref(noise), ea.ptr, ya.ptr,
                        out.ctypes.data, C.byref(st), ua.B)
    icnf.last_stats = _stats_from(st)
    icnf._check(rc)
    return out


# ---------------------------------------------------------------- flow API
def inference(icnf: ICNF, mode: Mode, xs, *args, eps=None, seed=None, tspan=None, sample_offset: int = 0, **sol):
    """``inference(icnf, mode, xs, [ys,] ps, st)`` -> ``(logp, (E, n, A))``
    (base_icnf.jl:406-424).  Keyword extras: ``eps`` supplies the probe matrix
    (D' x B) instead of drawing it; ``tspan`` overrides the (steered) span."""
    ys, ps, st = _split(args, 2, "inference")
    icnf._set_params(ps)
    dev = _on_device(xs, eps, ys)
    d = icnf.nvariables + icnf.naugments
    xa = _Arg(xs, icnf.nvariables, "xs", dev)
    ea = _Arg(eps, d, "eps", dev)
    ya = _Arg(ys, icnf.nconditions, "ys", dev)
    if icnf.nconditions and ya.ptr is None:
        raise TypeError("conditioned ICNF: pass ys")
    t0, t1 = tspan if tspan is not None else icnf.steer_tspan(mode)
    noise = icnf._noise(mode, eps, seed, sample_offset)
    solver = icnf._solver(sol)
    if dev:
        logp = torch.empty(xa.B, dtype=torch.float32, device=xa.keep.device)
        regs = torch.empty((xa.B, 3), dtype=torch.float32, device=xa.keep.device)
        ds = _dev_stats(icnf)
        icnf._check(lib.icnf_inference_dev(icnf._h, mode.code, C.byref(solver), t0, t1, xa.ptr, C.byref(noise), ea.ptr,
                                           ya.ptr, logp.data_ptr(), regs.data_ptr(), ds.data_ptr(), xa.B, _stream(icnf)))
        icnf.last_stats = _read_dev_stats(icnf, ds)
        r = regs.t()
        return logp, (r[0], r[1], r[2])
    logp = np.empty(xa.B, dtype=np.float32)
    regs = np.empty((3, xa.B), dtype=np.float32, order="F")
    stt = _lib.Stats()
    rc = lib.icnf_inference(icnf._h, mode.code, C.byref(solver), t0, t1, xa.ptr, C.byref(noise), ea.ptr, ya.ptr,
                            logp.ctypes.data, regs.ctypes.data, C.byref(stt), xa.B)
    icnf.last_stats = _stats_from(stt)
    icnf._check(rc)
    return logp, (regs[0], regs[1], regs[2])


def generate(icnf: ICNF, mode: Mode, *args, z0=None, eps=None, seed=None, tspan=None, sample_offset: int = 0, **sol):
    """``generate(icnf, mode, [ys,] ps, st, n)`` -> ``nvars x n`` samples
    (base_icnf.jl:351-404).  ``z0`` (D' x n) supplies the base sample instead of
    drawing it from N(0, I)."""
    ys, ps, st, n = _split(args, 3, "generate")
    icnf._set_params(ps)
    n = int(n)
    dev = _on_device(z0, eps, ys)
    d = icnf.nvariables + icnf.naugments
    za = _Arg(z0, d, "z0", dev)
    ea = _Arg(eps, d, "eps", dev)
    ya = _Arg(ys, icnf.nconditions, "ys", dev)
    if icnf.nconditions and (ya.ptr is None or ya.B != n):
        raise ValueError("conditioned ICNF: ys must have n columns (smoke_tests.jl:74,99)")
    if za.ptr is not None and za.B != n:
        raise ValueError("z0 must have n columns")
    t0, t1 = tspan if tspan is not None else icnf.steer_tspan(mode)
    noise = icnf._noise(mode, eps, seed, sample_offset)
    if noise.seed == 0 and z0 is None:
        noise.seed = int(icnf.rng.integers(0, 2 ** 63)) if seed is None else int(seed)
    solver = icnf._solver(sol)
    if dev:
        device = (za.keep if za.keep is not None else ya.keep if ya.keep is not None else ea.keep).device
        out = torch.empty((n, icnf.nvariables), dtype=torch.float32, device=device)
        ds = _dev_stats(icnf)
        icnf._check(lib.icnf_generate_dev(icnf._h, mode.code, C.byref(solver), t0, t1, za.ptr, C.byref(noise), ea.ptr,
                                          ya.ptr, out.data_ptr(), ds.data_ptr(), n, _stream(icnf)))
        icnf.last_stats = _read_dev_stats(icnf, ds)
        return out.t()
    out = np.empty((icnf.nvariables, n), dtype=np.float32, order="F")
    stt = _lib.Stats()
    rc = lib.icnf_generate(icnf._h, mode.code, C.byref(solver), t0, t1, za.ptr, C.byref(noise), ea.ptr, ya.ptr,
                           out.ctypes.data, C.byref(stt), n)
    icnf.last_stats = _stats_from(stt)
    icnf._check(rc)
    return out


def _loss_impl(icnf: ICNF, mode: Mode, xs, ys, ps, want_grad: bool, want_dxs: bool, eps, seed, tspan,
               sample_offset: int, global_batch: int, sol: dict, dp: bool = False):
    icnf._set_params(ps)
    dev = _on_device(xs, eps, ys)
    d = icnf.nvariables + icnf.naugments
    xa = _Arg(xs, icnf.nvariables, "xs", dev)
    ea = _Arg(eps, d, "eps", dev)
    ya = _Arg(ys, icnf.nconditions, "ys", dev)
    if icnf.nconditions and ya.ptr is None:
        raise TypeError("conditioned ICNF: pass ys")
    t0, t1 = tspan if tspan is not None else icnf.steer_tspan(mode)
    noise = icnf._noise(mode, eps, seed, sample_offset)
    solver = icnf._solver(sol)
    npar = icnf.n_params_lib
    planar = icnf.planar
    if dev:
        device = xa.keep.device
        # gradient and loss share one buffer [dtheta; loss] so that a data-parallel caller can
        # all-reduce both with a single collective and no concatenation
        buf = torch.empty(npar + 1, dtype=torch.float32, device=device)
        lossv = buf[npar:]
        dth = buf[:npar] if want_grad else None
        icnf._grad_loss_buf = buf if want_grad else None
        dxs = torch.empty((xa.B, icnf.nvariables), dtype=torch.float32, device=device) if want_dxs else None
        ds = _dev_stats(icnf)
        fn = lib.icnf_loss_grad_dp_dev if (dp and want_grad) else lib.icnf_loss_grad_dev
        icnf._check(fn(icnf._h, mode.code, C.byref(solver), t0, t1, xa.ptr, C.byref(noise), ea.ptr,
                       ya.ptr, lossv.data_ptr(), dth.data_ptr() if want_grad else None,
                       dxs.data_ptr() if want_dxs else None, ds.data_ptr(), xa.B, int(global_batch), _stream(icnf)))
        icnf._pending_stats = ds
        if planar is not None and dth is not None:
            dth = planar.grad_from_lib(dth)
        return lossv[0], dth, (dxs.t() if want_dxs else None)
    lossv = C.c_float()
    stt = _lib.Stats()
    if want_grad:
        dth = np.empty(npar, dtype=np.float32)
        dxs = np.empty((icnf.nvariables, xa.B), dtype=np.float32, order="F") if want_dxs else None
        fn = lib.icnf_loss_grad_dp if dp else lib.icnf_loss_grad
        rc = fn(icnf._h, mode.code, C.byref(solver), t0, t1, xa.ptr, C.byref(noise), ea.ptr, ya.ptr,
                C.byref(lossv), dth.ctypes.data, dxs.ctypes.data if want_dxs else None, C.byref(stt),
                xa.B, int(global_batch))
    else:
        dth = dxs = None
        rc = lib.icnf_loss(icnf._h, mode.code, C.byref(solver), t0, t1, xa.ptr, C.byref(noise), ea.ptr, ya.ptr,
                           C.byref(lossv), C.byref(stt), xa.B, int(global_batch))
    icnf.last_stats = _stats_from(stt)
    icnf._check(rc)
    if planar is not None and dth is not None:
        dth = planar.grad_from_lib(dth)
    return float(lossv.value), dth, dxs


def loss(icnf: ICNF, mode: Mode, xs, *args, eps=None, seed=None, tspan=None, sample_offset: int = 0,
         global_batch: int = 0, **sol):
    """``loss(icnf, mode, xs, [ys,] ps, st)`` (icnf.jl:628-649)."""
    ys, ps, st = _split(args, 2, "loss")
    return _loss_impl(icnf, mode, xs, ys, ps, False, False, eps, seed, tspan, sample_offset, global_batch, sol)[0]


def loss_and_gradient(icnf: ICNF, mode: Mode, xs, *args, want_dxs: bool = False, eps=None, seed=None, tspan=None,
                      sample_offset: int = 0, global_batch: int = 0, data_parallel: bool = False, **sol):
    """What ``Zygote.gradient(p -> loss(icnf, mode, xs, [ys,] p, st), ps)`` (and the
    gradient w.r.t. ``xs``, smoke_tests.jl:132-133) gives the reference's callers:
    returns ``(loss, dtheta)`` or ``(loss, dtheta, dxs)``.  ``data_parallel=True`` (a handle that joined a
    group, ``group_join``): ``xs`` is this rank's shard and the returned loss / gradient are those of the
    global batch, summed inside the library (``icnf_loss_grad_dp``)."""
    ys, ps, st = _split(args, 2, "loss_and_gradient")
    l, g, gx = _loss_impl(icnf, mode, xs, ys, ps, True, want_dxs, eps, seed, tspan, sample_offset, global_batch, sol,
                          dp=data_parallel)
    return (l, g, gx) if want_dxs else (l, g)


# ---------------------------------------------------------------- multi-GPU groups (include/icnf_b200.h)
def group_unique_id() -> bytes:
    buf = C.create_string_buffer(128)
    rc = lib.icnf_group_unique_id(buf)
    if rc != _lib.OK:
        raise ICNFError(rc, lib.icnf_last_error(None).decode())
    return buf.raw


def group_join_id(icnf: ICNF, uid: bytes, rank: int, world: int):
    """``icnf_group_join``: collective over the ``world`` processes that hold ``uid``."""
    buf = C.create_string_buffer(uid, 128)
    icnf._check(lib.icnf_group_join(icnf._h, buf, int(world), int(rank)))
    icnf._group = (int(rank), int(world))


def group_info(icnf: ICNF):
    n, r, p = C.c_int32(), C.c_int32(), C.c_int32()
    icnf._check(lib.icnf_group_info(icnf._h, C.byref(n), C.byref(r), C.byref(p)))
    return {"n_ranks": n.value, "rank": r.value, "peer_memory": bool(p.value)}


def group_set_global_norm(icnf: ICNF, enabled: bool = True):
    """``icnf_group_set_global_norm``: adaptive data-parallel solves use the error norm of the GLOBAL batch (one
    scalar pair per step attempt through NVLink peer memory), i.e. exactly the unsharded solve's steps."""
    icnf._check(lib.icnf_group_set_global_norm(icnf._h, int(bool(enabled))))


def create_group(icnfs: Sequence[ICNF]):
    """``icnf_create_group``: the handles of ONE process, one per device."""
    arr = (C.c_void_p * len(icnfs))(*[i._h.value for i in icnfs])
    rc = lib.icnf_create_group(arr, len(icnfs))
    if rc != _lib.OK:
        raise ICNFError(rc, lib.icnf_last_error(icnfs[0]._h).decode())
    for r, i in enumerate(icnfs):
        i._group = (r, len(icnfs))
