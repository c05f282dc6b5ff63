"""Timeline of CTA 0 of the last GEMM chain of a slot (development aid): ICNF_CHAIN_TRACE=<slot> python scripts/chain_timeline.py"""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import cnf_b200 as m
B = 8192
ffjord = m.Chain(m.Dense(785, 512, "softplus"), m.Dense(512, 512, "softplus"), m.Dense(512, 512, "softplus"), m.Dense(512, 784))
icnf = m.ICNF(nvariables=784, naugments=0, nn=ffjord, precision="bf16x3_tc", epsdist="rademacher")
rng = np.random.default_rng(7)
theta, _ = m.setup(rng, icnf)
xs = torch.from_numpy(rng.standard_normal((B, 784)).astype(np.float32)).cuda()
what = sys.argv[1] if len(sys.argv) > 1 else "fwd"
if what == "c5":   # config 5: 64-D conditioned flow at batch 65 536, TestMode (exact trace): 4 GEMMs per evaluation
    icnf = m.ICNF(nvariables=64, naugments=0, nconditions=32, precision="bf16x3_tc")
    theta, _ = m.setup(rng, icnf)
    xs = torch.from_numpy(rng.standard_normal((65536, 64)).astype(np.float32)).cuda()
    ys = torch.from_numpy(rng.standard_normal((65536, 32)).astype(np.float32)).cuda()
for i in range(2):
    if what == "c5":
        m.inference(icnf, m.TestMode(), xs.t(), ys.t(), theta, {}, adaptive=False, dt=0.25)
    elif what == "fwd":
        m.inference(icnf, m.TrainMode(False), xs.t(), theta, {}, seed=3, adaptive=False, dt=0.25)
    else:
        m.loss_and_gradient(icnf, m.TrainMode(True), xs.t(), torch.from_numpy(theta).cuda(), {}, seed=3, adaptive=False, dt=0.25)
torch.cuda.synchronize()
buf = np.zeros(3 * 8192, np.int64)
fn = m.lib.icnf_tc_chain_trace_fetch
fn.restype = C.c_int; fn.argtypes = [C.c_void_p]
assert fn(buf.ctypes.data) == 0
ev = []
for r in range(3):
    b = buf[r * 8192:(r + 1) * 8192]
    n = int(b[0])
    ev += [(int(b[2 + 2 * i]), int(b[1 + 2 * i])) for i in range(n)]
ev.sort()
t0 = ev[0][0]
names = {1: "tma", 3: "mma", 4: "acc_ready", 5: "epi_done"}
print(f"{len(ev)} events, span {ev[-1][0] - t0} cycles")
for t, tag in ev:
    if tag // 1000 in (4, 5) or "-v" in sys.argv:
        print(f"{t - t0:8d}  {names.get(tag // 1000, tag // 1000):10s} {tag % 1000}")

# summary of the MMA thread: cycles between consecutive K-block commits, and of the TMA thread between issues
for code, label in ((3, "MMA K-block commit"), (1, "TMA K-block issue")):
    ts = np.array([t for t, tag in ev if tag // 1000 == code])
    if len(ts) > 2:
        d = np.diff(ts)
        print(f"{label}: {len(ts)} events, interval median {np.median(d):.0f}, p10 {np.percentile(d, 10):.0f}, p90 {np.percentile(d, 90):.0f}, "
              f"mean {d.mean():.0f} cycles; intervals > 2x median: {(d > 2 * np.median(d)).sum()} holding {d[d > 2 * np.median(d)].sum() / d.sum():.0%} of the span")
