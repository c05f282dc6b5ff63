import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import cnf_b200 as m
from oracle import icnf_oracle as O
from tests.helpers import SHAPES, make_icnf, make_inputs, t64

icnf = make_icnf(m, "cond")
om, theta, xs, eps, ys = make_inputs(icnf, 200)
theta = (4.0 * theta).astype(np.float32)
mode = m.TrainMode(True)
bad = []
for b in range(200):
    try:
        m.inference(icnf, mode, xs[:, b:b+1], ys[:, b:b+1], theta, {}, eps=eps[:, b:b+1], tspan=icnf.tspan)
    except Exception as e:
        bad.append(b)
print("bad samples (B=1 adaptive):", bad)
# fixed-step sweep over all samples: where does the state go non-finite?
u0 = O.make_u0(om, t64(xs)).numpy().astype(np.float32)
for T in (0.5, 0.9, 0.94, 0.95, 1.0):
    u = m.base_sol(icnf, mode, u0, theta, tspan=(0.0, T), eps=eps, ys=ys, adaptive=False, dt=0.005)
    ref = O.solve(om, O.TRAIN_REG, t64(u0), t64(theta), t64(eps), t64(ys), 0.0, T, O.SolverOpts(adaptive=False, dt=0.005)).numpy()
    nf = np.argwhere(~np.isfinite(u))
    print("T", T, "nonfinite entries", nf[:5].tolist(), "max|u|", np.nanmax(np.abs(u)), "max err vs oracle", np.nanmax(np.abs(u - ref)))
# half batches
for lo, hi in [(0, 100), (100, 200), (0, 200)]:
    try:
        m.inference(icnf, mode, xs[:, lo:hi], ys[:, lo:hi], theta, {}, eps=eps[:, lo:hi], tspan=icnf.tspan)
        print("range", lo, hi, "ok", icnf.last_stats)
    except Exception as e:
        print("range", lo, hi, "FAIL", icnf.last_stats)
