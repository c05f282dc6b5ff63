import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import cnf_b200 as m
B = 65536
icnf = m.ICNF(nvariables=2, naugments=0, epsdist="rademacher")
rng = np.random.default_rng(0)
theta, _ = m.setup(rng, icnf)
xs = rng.standard_normal((2, B)).astype(np.float32)
icnf.set_profiling(True)
def run(label, mode, **kw):
    ts = []
    for i in range(6):
        m.inference(icnf, mode, xs, theta, {}, seed=i, tspan=(0.0, 1.0), **kw)
        ts.append(icnf.kernel_times_ms()["forward"])
    st = icnf.last_stats
    print(f"{label:40s} fwd {np.median(ts[2:])*1e3:8.1f} us  steps {st.naccept}+{st.nreject} nf {st.nf}", flush=True)
for mode in (m.TrainMode(True), m.TestMode()):
    print(mode)
    run("adaptive tol 1e-4", mode)
    run("adaptive tol 1e-4, dt0 given", mode, dt=0.1)
    run("adaptive tol 1e-2", mode, reltol=1e-2, abstol=1e-2)
    run("adaptive tol 1e-6", mode, reltol=1e-6, abstol=1e-6)
    run("fixed 5 steps", mode, adaptive=False, dt=0.2)
    run("fixed 10 steps", mode, adaptive=False, dt=0.1)
    run("fixed 20 steps", mode, adaptive=False, dt=0.05)
