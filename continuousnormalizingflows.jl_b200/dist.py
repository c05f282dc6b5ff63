"""``ICNFDist`` / ``CondICNFDist``: the Distributions.jl adapter of the reference
(/root/reference/src/exts/dist_ext/core_icnf.jl:1-75, core_cond_icnf.jl) as a
thin caller of the B200 flow API: ``logpdf`` -> ``inference``, ``rand`` ->
``generate``."""

from __future__ import annotations

import numpy as np

from .api import ICNF, Mode, generate, inference, _is_torch


class ICNFDist:
    def __init__(self, icnf: ICNF, mode: Mode, ps, st):
        self.icnf, self.mode, self.ps, self.st = icnf, mode, ps, st

    def __len__(self):                      # Base.length (dist_ext/core.jl:6-8)
        return self.icnf.nvariables

    def _cond(self, n):
        return ()

    def logpdf(self, x):
        """``Distributions._logpdf`` (core_icnf.jl:36-41): x is nvars x B, or a length-nvars vector."""
        vec = getattr(x, "ndim", 2) == 1
        if vec:
            x = x.reshape(self.icnf.nvariables, 1)
        out = inference(self.icnf, self.mode, x, *self._cond(x.shape[1]), self.ps, self.st)[0]
        return out[0] if vec else out

    def pdf(self, x):
        lp = self.logpdf(x)
        return lp.exp() if _is_torch(lp) else np.exp(lp)

    def rand(self, n=None):
        """``Distributions._rand!`` (core_icnf.jl:69-75): nvars x n samples (a vector when n is None)."""
        m = 1 if n is None else int(n)
        out = generate(self.icnf, self.mode, *self._cond(m), self.ps, self.st, m)
        return out[:, 0] if n is None else out


class CondICNFDist(ICNFDist):
    def __init__(self, icnf: ICNF, mode: Mode, ys, ps, st):
        super().__init__(icnf, mode, ps, st)
        self.ys = ys

    def _cond(self, n):
        ys = self.ys
        if ys.shape[1] != n:
            raise ValueError("ys must have one column per sample")
        return (ys,)
