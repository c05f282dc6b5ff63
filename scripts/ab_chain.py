"""A/B of the tensor-core path's tuning knobs inside ONE process (alternating rounds, medians): box-to-box and
run-to-run variation is larger than most of the effects being measured."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import cnf_b200 as m
what = sys.argv[1] if len(sys.argv) > 1 else "c4"
rounds = int(sys.argv[2]) if len(sys.argv) > 2 else 8
kw = dict(adaptive=False, dt=0.25) if (len(sys.argv) <= 3 or sys.argv[3] != "adaptive") else {}
rng = np.random.default_rng(7)
if what == "c4":
    B = 8192
    net = m.Chain(m.Dense(785, 512, "softplus"), m.Dense(512, 512, "softplus"), m.Dense(512, 512, "softplus"), m.Dense(512, 784))
    icnf = m.ICNF(nvariables=784, naugments=0, nn=net, precision="bf16x3_tc", epsdist="rademacher")
    xs = torch.from_numpy(rng.standard_normal((B, 784)).astype(np.float32)).cuda()
    ys = None
    mode_inf = m.TrainMode(False)
elif what == "c5":
    B = 65536
    icnf = m.ICNF(nvariables=64, naugments=0, nconditions=32, precision="bf16x3_tc")
    xs = torch.from_numpy(rng.standard_normal((B, 64)).astype(np.float32)).cuda()
    ys = torch.from_numpy(rng.standard_normal((B, 32)).astype(np.float32)).cuda()
    mode_inf = m.TestMode()
else:
    B = 262144
    icnf = m.ICNF(nvariables=16, naugments=0, precision="bf16x3_tc")
    xs = torch.from_numpy(rng.standard_normal((B, 16)).astype(np.float32)).cuda()
    ys = None
    mode_inf = m.TestMode()
theta, _ = m.setup(rng, icnf)
theta_d = torch.from_numpy(theta).cuda()
knob = m.lib.icnf_tc_knob_set
variants = {"unchained": [(0, 0), (2, 0), (3, 1), (4, 2)], "chain": [(0, 1), (1, 0), (2, 0), (3, 1), (4, 2)],
            "chain, 4 accumulators": [(0, 1), (1, 0), (2, 0), (3, 1), (4, 4)],
            "chain, super-groups (3 waves)": [(0, 1), (1, 0), (2, 1), (3, 1), (4, 2)],
            "chain cluster 2": [(0, 1), (1, 0), (2, 0), (3, 2), (4, 2)]}
def timed(fn):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); fn(); b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b)
def inf():
    if ys is None: m.inference(icnf, mode_inf, xs.t(), theta, {}, seed=3, **kw)
    else: m.inference(icnf, mode_inf, xs.t(), ys.t(), theta, {}, seed=3, **kw)
def train():
    m.loss_and_gradient(icnf, m.TrainMode(True), xs.t(), theta_d, {}, seed=3, **kw)
res = {k: {"inf": [], "train": []} for k in variants}
for r in range(rounds + 1):
    for name, ks in variants.items():
        for k, v in ks: knob(k, v)
        ti = timed(inf)
        tt = timed(train) if what == "c4" else 0.0
        if r > 0:
            res[name]["inf"].append(ti); res[name]["train"].append(tt)
for name in variants:
    print(f"{what} {name:30s} inference 24 RHS: median {np.median(res[name]['inf']):7.3f} min {np.min(res[name]['inf']):7.3f} ms"
          + (f"   training step (4 fixed steps): median {np.median(res[name]['train']):7.3f} min {np.min(res[name]['train']):7.3f} ms" if what == "c4" else ""), flush=True)
