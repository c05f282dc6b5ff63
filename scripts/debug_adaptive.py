import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import cnf_b200 as m
from oracle import icnf_oracle as O
from tests.helpers import SHAPES, make_icnf, make_inputs, t64

for shape, scale in [("config2_moons", 1.0), ("cond", 1.0), ("config1_usage", 4.0), ("cond", 4.0)]:
    icnf = make_icnf(m, shape)
    om, theta, xs, eps, ys = make_inputs(icnf, 200)
    theta = (scale * theta).astype(np.float32)
    for mode, omode in [(m.TestMode(), O.TEST), (m.TrainMode(True), O.TRAIN_REG), (m.TrainMode(False), O.TRAIN_NOREG)]:
        args = (xs,) if ys is None else (xs, ys)
        st = O.SolveStats()
        rl, (rE, rn, rA) = O.inference(om, omode, t64(xs), t64(theta), t64(eps), t64(ys), stats=st)
        try:
            logp, (E, n, A) = m.inference(icnf, mode, *args, theta, {}, eps=eps, tspan=icnf.tspan)
            gs = icnf.last_stats
            err = lambda a, b: float(np.max(np.abs(a - b.numpy()) / (np.abs(b.numpy()) + 1e-5)))
            print(shape, scale, mode, "gpu", gs.naccept, gs.nreject, gs.nf, "oracle", st.naccept, st.nreject, st.nf,
                  "err logp %.2e E %.2e n %.2e A %.2e" % (err(logp, rl), err(E, rE), err(n, rn), err(A, rA)), flush=True)
        except Exception as e:
            print(shape, scale, mode, "FAILED", e, icnf.last_stats, "oracle", st.naccept, st.nreject, [round(x, 4) for x in st.dts[:8]], flush=True)
