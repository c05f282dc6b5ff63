"""``ICNFModel`` / ``CondICNFModel``: the MLJ adapter of the reference
(/root/reference/src/exts/mlj_ext/core_icnf.jl:1-95, core.jl:1-105) as host
orchestration around the B200 loss/gradient call: minibatches with
``shuffle=true, partial=true`` (core.jl:24-35), ``OptimiserChain(WeightDecay(1e-4),
Adam(1e-3, (0.9, 0.999), 1e-8))`` (core_icnf.jl:17-24), ``epochs = 300`` (:15).
``transform`` returns ``exp.(logp)`` in ``TestMode`` (core_icnf.jl:60-68)."""

from __future__ import annotations

from dataclasses import dataclass
from typing import Callable, Optional

import numpy as np

from .api import ICNF, TestMode, TrainMode, inference, loss_and_gradient, setup


@dataclass
class Adam:
    eta: float = 1e-3
    beta: tuple = (0.9, 0.999)
    epsilon: float = 1e-8


@dataclass
class WeightDecay:
    lam: float = 1e-4


class _DeviceOptimiserChain:
    """The same optimiser with theta, m and v resident on the GPU (icnf_adam_step_dev): the
    gradient never leaves the device between icnf_loss_grad_dev and the update."""

    def __init__(self, wd: WeightDecay, adam: Adam, theta0: np.ndarray, device: int):
        import torch
        self.torch = torch
        self.wd, self.adam = wd, adam
        self.theta = torch.tensor(theta0, dtype=torch.float32, device=f"cuda:{device}")
        self.m = torch.zeros_like(self.theta)
        self.v = torch.zeros_like(self.theta)
        self.t = 0

    def step(self, grad) -> None:
        from ._lib import lib
        self.t += 1
        b1, b2 = self.adam.beta
        rc = lib.icnf_adam_step_dev(self.theta.data_ptr(), grad.data_ptr(), self.m.data_ptr(), self.v.data_ptr(),
                                    self.theta.numel(), self.t, self.adam.eta, b1, b2, self.adam.epsilon, self.wd.lam,
                                    self.torch.cuda.current_stream(self.theta.device).cuda_stream)
        if rc != 0:
            raise RuntimeError(f"icnf_adam_step_dev failed with status {rc}")


class _OptimiserChain:
    """Optimisers.jl semantics: WeightDecay adds lambda*theta to the gradient, Adam
    then rescales it; theta <- theta - update."""

    def __init__(self, wd: WeightDecay, adam: Adam, n: int):
        self.wd, self.adam = wd, adam
        self.m = np.zeros(n, np.float32)
        self.v = np.zeros(n, np.float32)
        self.bp = np.array([1.0, 1.0])

    def step(self, theta: np.ndarray, grad: np.ndarray) -> np.ndarray:
        g = grad + np.float32(self.wd.lam) * theta
        b1, b2 = self.adam.beta
        self.m = b1 * self.m + (1 - b1) * g
        self.v = b2 * self.v + (1 - b2) * g * g
        self.bp *= (b1, b2)
        upd = self.m / (1 - self.bp[0]) / (np.sqrt(self.v / (1 - self.bp[1])) + self.adam.epsilon) * self.adam.eta
        return (theta - upd).astype(np.float32)


def make_opt_callback(n: int) -> Callable:
    def cb(it: int, l: float) -> bool:     # core.jl:96-105
        if it % n == 1:
            print(f"Iteration: {it} | Loss: {l}")
        return False
    return cb


class ICNFModel:
    def __init__(self, icnf: Optional[ICNF] = None, batchsize: int = 1024, epochs: int = 300,
                 adam: Adam = Adam(), weight_decay: WeightDecay = WeightDecay(), callback: Optional[Callable] = None,
                 rng=None, device_optimiser: bool = False):
        self.icnf = icnf if icnf is not None else ICNF()
        self.batchsize, self.epochs = int(batchsize), int(epochs)
        self.adam, self.weight_decay = adam, weight_decay
        self.callback = callback
        self.device_optimiser = bool(device_optimiser)   # keep data, parameters and optimiser state on the GPU
        self.rng = np.random.default_rng(rng)
        self.fitresult = None
        self.report = {}

    def _batches(self, n: int):
        bs = n if self.batchsize == 0 else self.batchsize      # get_batchsize, core.jl:37-43
        perm = self.rng.permutation(n)                         # shuffle = true
        for i in range(0, n, bs):                              # partial = true
            yield perm[i:i + bs]

    def fit(self, X, Y=None):
        """X: table as an (n_samples, nvars) array (MLJ tables are row = observation)."""
        x = np.ascontiguousarray(np.asarray(X, dtype=np.float32).T)       # permutedims(matrix(X)), core_icnf.jl:33
        y = None if Y is None else np.ascontiguousarray(np.asarray(Y, dtype=np.float32).T)
        ps, st = setup(self.icnf.rng, self.icnf)
        it, last, stop = 0, float("nan"), False
        if self.device_optimiser:
            import torch
            dev = f"cuda:{self.icnf.device}"
            opt = _DeviceOptimiserChain(self.weight_decay, self.adam, ps, self.icnf.device)
            xd = torch.tensor(x.T.copy(), device=dev)                      # (n, nvars) records
            yd = None if y is None else torch.tensor(y.T.copy(), device=dev)
            l = None
            for _ in range(self.epochs):
                for idx in self._batches(x.shape[1]):
                    it += 1
                    sel = torch.as_tensor(idx, device=dev)
                    args = (xd[sel].t(),) if yd is None else (xd[sel].t(), yd[sel].t())
                    l, g = loss_and_gradient(self.icnf, TrainMode(True), *args, opt.theta, st)
                    opt.step(g)
                    if self.callback:
                        last = float(l)
                        self.icnf.check_last()           # the loss was read anyway: the stream is already drained
                        stop = bool(self.callback(it, last))
                    elif it % 64 == 0:
                        self.icnf.check_last()           # a failed solve returns a NaN gradient: stop instead of training on it
                    if stop:
                        break
                if stop:                                 # the reference's callback halts the whole optimisation (core.jl:96-105)
                    break
            if l is not None:
                last = float(l)
                self.icnf.check_last()
            ps = opt.theta.cpu().numpy()
        else:
            opt = _OptimiserChain(self.weight_decay, self.adam, ps.size)
            for _ in range(self.epochs):
                for idx in self._batches(x.shape[1]):
                    it += 1
                    args = (x[:, idx],) if y is None else (x[:, idx], y[:, idx])
                    last, g = loss_and_gradient(self.icnf, TrainMode(True), *args, ps, st)
                    ps = opt.step(ps, g)
                    if self.callback and self.callback(it, last):
                        stop = True
                        break
                if stop:
                    break
        self.fitresult = (ps, st)
        self.report = {"iterations": it, "final_loss": last}
        return self

    def fitted_params(self):
        ps, st = self.fitresult
        return {"learned_parameters": ps, "states": st}

    def transform(self, Xnew, Ynew=None):
        x = np.ascontiguousarray(np.asarray(Xnew, dtype=np.float32).T)
        ps, st = self.fitresult
        args = (x,) if Ynew is None else (x, np.ascontiguousarray(np.asarray(Ynew, dtype=np.float32).T))
        logp = inference(self.icnf, TestMode(), *args, ps, st)[0]
        return {"px": np.exp(logp)}


class CondICNFModel(ICNFModel):
    pass
