"""Shared helpers for the parity tests: build the oracle's view of a product
``ICNF`` and seeded inputs in the reference's R x B layout."""
import numpy as np
import torch

from oracle import icnf_oracle as O

ACT_CODE = {"softplus": O.ACT_SOFTPLUS, "tanh": O.ACT_TANH, "sigmoid": O.ACT_SIGMOID, "identity": O.ACT_IDENTITY}

# (kwargs for cnf_b200.ICNF, hidden activation) -- every shape has a tiny-family instantiation
SHAPES = {
    "config1_usage": dict(nvariables=1),                                             # 4-16-16-3 softplus
    "config2_moons": dict(nvariables=2, naugments=0),                                # 3-12-12-2 softplus
    "smoke_default": dict(nvariables=2),                                             # 6-24-24-5 softplus
    "cond": dict(nvariables=2, naugments=1, nconditions=2, n_hidden=8),              # 6-8-8-3 softplus
    "tanh_auto_1hidden": dict(nvariables=3, naugments=1, autonomous=True, nn=("tanh", (4, 8, 4))),
    "sigmoid_3hidden": dict(nvariables=2, naugments=0, nn=("sigmoid", (3, 7, 9, 5, 2))),
}


# shapes served by the generic family (any width; fp32 tiled SGEMMs)
GENERIC_SHAPES = {
    "config3_gmm16": dict(nvariables=16, naugments=0),                                      # 17-68-68-16 softplus
    "cond_generic": dict(nvariables=6, naugments=1, nconditions=3, n_hidden=40),           # 11-40-40-7 softplus
    "deep4_tanh": dict(nvariables=5, naugments=0, nn=("tanh", (6, 33, 47, 21, 5))),         # exact trace by one-hot chains
    "one_hidden_sigmoid": dict(nvariables=7, naugments=0, autonomous=True, nn=("sigmoid", (7, 50, 7))),
    "linear9": dict(nvariables=9, naugments=0, autonomous=True, nn=("identity", (9, 9))),
}


def make_icnf(m, name, **extra):
    kw = dict(SHAPES[name] if name in SHAPES else GENERIC_SHAPES[name])
    nn = kw.pop("nn", None)
    if nn is not None:
        act, sizes = nn
        layers = [m.Dense(sizes[i], sizes[i + 1], act if i < len(sizes) - 2 else "identity") for i in range(len(sizes) - 1)]
        kw["nn"] = m.Chain(*layers)
    kw.update(extra)
    return m.ICNF(**kw)


def oracle_model(icnf):
    acts = {l.activation for l in icnf.nn.layers[:-1]} or {"identity"}
    return O.OracleICNF(nvars=icnf.nvariables, naug=icnf.naugments, ncond=icnf.nconditions,
                        autonomous=icnf.autonomous, hidden=tuple(icnf.sizes[1:-1]), activation=ACT_CODE[acts.pop()],
                        lam1=icnf.lambda1, lam2=icnf.lambda2, lam3=icnf.lambda3, tspan=icnf.tspan,
                        steer_rate=icnf.steer_rate)


def make_inputs(icnf, B, seed=0, bias_scale=0.3):
    om = oracle_model(icnf)
    rng = np.random.default_rng(seed)
    theta = O.init_params(om, seed + 1, np.float32, bias_scale=bias_scale)
    xs = rng.standard_normal((icnf.nvariables, B)).astype(np.float32)
    eps = rng.standard_normal((om.d, B)).astype(np.float32)
    ys = rng.standard_normal((icnf.nconditions, B)).astype(np.float32) if icnf.nconditions else None
    return om, theta, xs, eps, ys


def t64(a):
    return None if a is None else torch.tensor(np.asarray(a), dtype=torch.float64)


def rel_err(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.max(np.abs(a - b) / (np.abs(b) + 1e-6))) if a.size else 0.0


def norm_rel_err(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.linalg.norm(a - b) / (np.linalg.norm(b) + 1e-30))
