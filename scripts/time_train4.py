import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import cnf_b200 as m
prec = sys.argv[1] if len(sys.argv) > 1 else "fp32"
B = 8192
ffjord = m.Chain(m.Dense(785, 512, "softplus"), m.Dense(512, 512, "softplus"), m.Dense(512, 512, "softplus"), m.Dense(512, 784))
icnf = m.ICNF(nvariables=784, naugments=0, nn=ffjord, precision=prec, epsdist="rademacher")
rng = np.random.default_rng(7)
theta, _ = m.setup(rng, icnf)
theta_d = torch.from_numpy(theta).cuda()
xs = torch.from_numpy(rng.standard_normal((B, 784)).astype(np.float32)).cuda()
for i in range(6):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    m.loss_and_gradient(icnf, m.TrainMode(True), xs.t(), theta_d, {}, seed=3, adaptive=False, dt=0.25)
    b.record(); torch.cuda.synchronize()
    print(prec, "config 4 training step, 4 fixed steps:", a.elapsed_time(b), "ms", os.environ.get("ICNF_TC_MODE", "auto"), flush=True)
