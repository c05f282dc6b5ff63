#!/usr/bin/env python
"""Generates tests/golden/icnf_golden.npz from the float64 oracle (oracle/icnf_oracle.py).

These are NOT outputs of the Julia reference (it cannot run here: parity is unpinned, see
DESIGN.md); they pin the oracle itself against accidental change and give the CUDA path a
fixed set of vectors that does not depend on importing the oracle at test time.

    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import icnf_oracle as O  # noqa: E402

CASES = {
    # name: (model kwargs, batch)
    "usage_1d": (dict(nvars=1), 16),                                   # examples/usage.jl shape 4-16-16-3
    "moons_2d": (dict(nvars=2, naug=0), 16),                           # config 2 shape 3-12-12-2
    "cond_2d": (dict(nvars=2, naug=1, ncond=2, hidden=(8, 8)), 12),
    "gmm_16d": (dict(nvars=16, naug=0), 8),                            # config 3 shape 17-68-68-16
}


def main():
    out = {}
    t = lambda a: None if a is None else torch.tensor(a, dtype=torch.float64)
    for name, (kw, B) in CASES.items():
        om = O.OracleICNF(**kw)
        rng = np.random.default_rng(abs(hash(name)) % 2 ** 31 if False else sum(map(ord, name)))
        theta = O.init_params(om, 7, np.float32, bias_scale=0.25)
        xs = rng.standard_normal((om.nvars, B)).astype(np.float32)
        eps = rng.standard_normal((om.d, B)).astype(np.float32)
        ys = rng.standard_normal((om.ncond, B)).astype(np.float32) if om.ncond else None
        u = rng.standard_normal((om.n_state, B)).astype(np.float32)
        out[f"{name}/theta"], out[f"{name}/xs"], out[f"{name}/eps"], out[f"{name}/u"] = theta, xs, eps, u
        if ys is not None:
            out[f"{name}/ys"] = ys
        opts = O.SolverOpts(adaptive=False, dt=0.125)
        for mode, tag in ((O.TEST, "test"), (O.TRAIN_REG, "train")):
            out[f"{name}/rhs_{tag}"] = O.rhs_closed(om, mode, t(u), t(theta), 0.37, t(eps), t(ys)).numpy()
            logp, (E, n, A) = O.inference(om, mode, t(xs), t(theta), t(eps), t(ys), opts=opts)
            out[f"{name}/logp_{tag}"] = logp.numpy()
            out[f"{name}/regs_{tag}"] = np.stack([E.numpy(), n.numpy(), A.numpy()])
        val, g, gx = O.loss_grad(om, O.TRAIN_REG, t(xs), t(theta), t(eps), t(ys), opts=opts, want_dxs=True)
        out[f"{name}/loss_train"] = np.array(float(val))
        out[f"{name}/dtheta_train"] = g.numpy()
        out[f"{name}/dxs_train"] = gx.numpy()
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "icnf_golden.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes,", len(out), "arrays")


if __name__ == "__main__":
    main()
