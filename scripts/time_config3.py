import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import cnf_b200 as m
prec = sys.argv[1] if len(sys.argv) > 1 else "fp32"
B = 262144
icnf = m.ICNF(nvariables=16, naugments=0, precision=prec)
rng = np.random.default_rng(7)
theta, _ = m.setup(rng, icnf)
xs = torch.from_numpy(rng.standard_normal((B, 16)).astype(np.float32)).cuda()
for i in range(6):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    m.inference(icnf, m.TestMode(), xs.t(), theta, {}, adaptive=False, dt=0.25)
    b.record(); torch.cuda.synchronize()
    print(prec, "config 3 TestMode fixed 4 steps (24 RHS):", a.elapsed_time(b), "ms", flush=True)
