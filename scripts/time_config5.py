import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import cnf_b200 as m
prec = sys.argv[1] if len(sys.argv) > 1 else "bf16x3_tc"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 65536
icnf = m.ICNF(nvariables=64, naugments=0, nconditions=32, precision=prec)
rng = np.random.default_rng(7)
theta, _ = m.setup(rng, icnf)
xs = torch.from_numpy(rng.standard_normal((B, 64)).astype(np.float32)).cuda()
ys = torch.from_numpy(rng.standard_normal((B, 32)).astype(np.float32)).cuda()
for i in range(6):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    m.inference(icnf, m.TestMode(), xs.t(), ys.t(), theta, {}, adaptive=False, dt=0.25)
    b.record(); torch.cuda.synchronize()
    print(prec, f"config 5 B={B} TestMode fixed 4 steps (24 RHS):", a.elapsed_time(b), "ms", os.environ.get("ICNF_CHAIN_SG", "sg"), os.environ.get("ICNF_TC_CHAIN", "chain"), flush=True)
