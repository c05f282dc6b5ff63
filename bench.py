#!/usr/bin/env python
"""bench.py -- headline measurement of the B200 hot path.

Workload (BASELINE.json configs[1]): 2-D two-moons RNODE, nvars 2, naug 0,
3 -> 12 -> 12 -> 2 softplus MLP (the reference's default width 4 n_in), one
Hutchinson Rademacher probe drawn in-kernel, regularisers lambda1 = lambda2 = 0.01,
STEER end time, Tsit5 adaptive at reltol = abstol = 1e-4 (the reference's default
tolerances), batch 65 536 per GPU.  A "step" is one training step: loss + gradient
w.r.t. all parameters (+ one all-reduce of [gradient; loss] when N > 1).

    python bench.py --gpus N --steps K --warmup W            # our arm
    python bench.py --impl reference --gpus N --steps K ...  # CPU restatement arm

Prints ONE JSON line (rank 0).  `value` = samples/s with the batch resident in
HBM; `e2e` = the same step through the host-pointer C ABI (pinned host xs in,
loss + gradient out) -- see DESIGN.md "Measurement".  `extras` carries the other
BASELINE.json configurations (log p(x) evaluations/s of configs 1-5, config-4 training
at every N, config-5 generate / inference over batch sizes); every extras case is first
checked against the CPU oracle on a sub-batch (the oracle is the checker, never the thing timed).
"""
import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np
import torch

WORKLOAD = "two-moons RNODE training step (config 2): nvars=2 naug=0 mlp=3-12-12-2 softplus, Hutchinson Rademacher in-kernel, lambda1=lambda2=0.01, STEER 0.1, Tsit5 adaptive rtol=atol=1e-4"
BATCH_PER_GPU = 65536
METRIC = "rnode_train_samples_per_sec"
UNIT = "samples/s"
NPARAMS = 3 * 12 + 12 + 12 * 12 + 12 + 12 * 2 + 2     # 230
PW = 3 * 12 + 12 * 12 + 12 * 2                         # weights only (P in SURVEY 8(d)) = 204


def make_config(batch, world):
    """identical keys in both arms, so that the driver can compare the two lines' configs"""
    return {"workload": WORKLOAD, "batch_per_gpu": batch, "global_batch": batch * world, "parallelism": f"dp{world}"}


def two_moons(n, seed=0):
    """Two interleaved half-circles of radius 1, offset (1, -0.5), N(0, 0.1^2) noise."""
    rng = np.random.default_rng(seed)
    k = n // 2
    a = rng.uniform(0, math.pi, size=n)
    x = np.where(np.arange(n) < k, np.cos(a), 1.0 - np.cos(a))
    y = np.where(np.arange(n) < k, np.sin(a), -0.5 - np.sin(a) + 1.0)
    pts = np.stack([x, y]) + 0.1 * rng.standard_normal((2, n))
    perm = rng.permutation(n)
    return np.ascontiguousarray(pts[:, perm]).astype(np.float32)


def init_theta(seed=0):
    """Lux Dense default init: glorot-uniform weights, zero bias (random-init weights, no checkpoints offline)."""
    rng = np.random.default_rng(seed)
    parts = []
    for nin, nout in ((3, 12), (12, 12), (12, 2)):
        lim = math.sqrt(6.0 / (nin + nout))
        parts += [rng.uniform(-lim, lim, size=nin * nout), np.zeros(nout)]
    return np.concatenate(parts).astype(np.float32)


# ------------------------------------------------------------------ clocks
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device):
        self.device, self.lines, self.proc = device, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(self.device)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


# ------------------------------------------------------------------ CPU restatement (oracle) arm
def cpu_loss_grad_rate(B, steps, warmup, threads):
    """Times the oracle's training step (loss + AD gradient through the discrete solve,
    the CPU restatement of the reference path) on `threads` host threads."""
    from oracle import icnf_oracle as O
    from oracle import philox as P
    torch.set_num_threads(threads)
    om = O.OracleICNF(nvars=2, naug=0)
    theta = torch.tensor(init_theta())
    xs = torch.tensor(two_moons(B, seed=1))
    times, nf = [], 0
    for i in range(warmup + steps):
        eps = torch.tensor(P.rademacher(100 + i, 2, B))
        t1 = O.steer_t1(om, O.TRAIN_REG, P.uniform_pm(100 + i, 0.1))
        st = O.SolveStats()
        t0 = time.perf_counter()
        O.loss_grad(om, O.TRAIN_REG, xs, theta, eps, t1=t1, stats=st)
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
            nf = st.nf
    return B * len(times) / sum(times), 1e3 * sum(times) / len(times), nf


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    B = args.cpu_batch
    rate, ms, nf = cpu_loss_grad_rate(B, args.steps, max(args.warmup, 1), threads)
    cfg = make_config(args.batch, args.gpus)
    line = {
        "impl": "reference", "metric": METRIC, "value": rate, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic", "config": cfg,
        "note": "CPU restatement of the reference path (oracle/, torch CPU fp32 + autograd); the Julia package cannot run here",
        "cpu_baseline": {"value": rate, "unit": UNIT, "cores": threads, "kind": "port",
                         "sample": f"{args.steps} training steps of a {B}-sample batch of the same workload (nf={nf} RHS calls per solve)"},
        "e2e": {"value": rate, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------ parity gate of the side measurements
def _oracle_of(icnf):
    from oracle import icnf_oracle as O
    acts = {"softplus": O.ACT_SOFTPLUS, "tanh": O.ACT_TANH, "sigmoid": O.ACT_SIGMOID, "identity": O.ACT_IDENTITY}
    a = {l.activation for l in icnf.nn.layers[:-1]} or {"identity"}
    return O.OracleICNF(nvars=icnf.nvariables, naug=icnf.naugments, ncond=icnf.nconditions, autonomous=icnf.autonomous,
                        hidden=tuple(icnf.sizes[1:-1]), activation=acts[a.pop()], lam1=icnf.lambda1, lam2=icnf.lambda2,
                        lam3=icnf.lambda3, tspan=icnf.tspan, steer_rate=icnf.steer_rate)


def _nrm(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.linalg.norm(a - b) / (np.linalg.norm(b) + 1e-30))


def parity_gate(m, icnf, mode, theta, what, sol, nb=192, seed=5):
    """Before a case is timed: the same handle, on a small seeded sub-batch with a SUPPLIED probe, against the float64
    oracle.  Returns the normalised error; raises if it is outside the precision's tolerance."""
    from oracle import icnf_oracle as O
    from oracle import philox as P
    om = _oracle_of(icnf)
    rng = np.random.default_rng(seed)
    xs = rng.standard_normal((icnf.nvariables, nb)).astype(np.float32)
    ys = rng.standard_normal((icnf.nconditions, nb)).astype(np.float32) if icnf.nconditions else None
    eps = P.rademacher(seed, om.d, nb).astype(np.float32)
    t = lambda x: None if x is None else torch.tensor(np.asarray(x), dtype=torch.float64)
    omode = O.TEST if isinstance(mode, m.TestMode) else (O.TRAIN_REG if mode.reg else O.TRAIN_NOREG)
    vcabm = str(sol.get("alg", "Tsit5")).lower().startswith("vcabm")
    opts = O.SolverOpts(alg="vcabm" if vcabm else "tsit5", adaptive=bool(sol.get("adaptive", True)), dt=float(sol.get("dt", 0.0)))
    ya = (ys,) if ys is not None else ()
    prec = getattr(icnf, "precision_name", "fp32")
    tol = 2e-2 if prec == "bf16_tc" else 1e-4
    if vcabm:
        tol = 2e-3      # float32 rounding may move an order / accept decision of the Adams stepper: agreement at solver tolerance
    if what == "inference":
        got, _ = m.inference(icnf, mode, xs, *ya, theta, {}, eps=eps, tspan=icnf.tspan, **sol)
        ref, _ = O.inference(om, omode, t(xs), t(theta), t(eps), t(ys), opts=opts)
        err = _nrm(got, ref.numpy())
    elif what == "generate":
        z0 = rng.standard_normal((om.d, nb)).astype(np.float32)
        got = m.generate(icnf, mode, *ya, theta, {}, nb, z0=z0, eps=eps, tspan=icnf.tspan, **sol)
        ref = O.generate(om, omode, t(z0), t(theta), t(eps), t(ys), opts=opts)
        err = _nrm(got, ref.numpy())
    else:   # training step: loss + gradient, fixed steps so that the discrete solves are the same object
        fx = dict(adaptive=False, dt=0.5)
        l, g = m.loss_and_gradient(icnf, mode, xs, *ya, theta, {}, eps=eps, tspan=icnf.tspan, **fx)
        rl, rg, _ = O.loss_grad(om, omode, t(xs), t(theta), t(eps), t(ys), opts=O.SolverOpts(adaptive=False, dt=0.5))
        err = max(_nrm(g, rg.numpy()), abs(l - float(rl)) / abs(float(rl)))
        tol = 3e-2 if prec == "bf16_tc" else 2e-4
    if not (err < tol):
        raise AssertionError(f"parity gate failed for {what} ({prec}): normalised error {err:.3e} >= {tol:.1e}")
    return err


def _timed_calls(fn, flush_buf, K=3, W=2):
    for _ in range(W):
        fn()
    torch.cuda.synchronize()
    ms = 0.0
    for _ in range(K):
        flush_buf.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ms += a.elapsed_time(b)
    return ms / K


def ffjord_chain(m):
    return m.Chain(m.Dense(785, 512, "softplus"), m.Dense(512, 512, "softplus"), m.Dense(512, 512, "softplus"), m.Dense(512, 784))


def _icnf(m, local, prec="fp32", **kw):
    icnf = m.ICNF(device=local, epsdist="rademacher", precision=prec, **kw)
    icnf.precision_name = prec
    return icnf


# ------------------------------------------------------------------ log p(x) side measurements (N = 1)
def logp_extras(m, local, dev, flush_buf):
    """The other half of BASELINE.json's metric: log-density evaluations per second
    (TestMode = exact trace, Tsit5 adaptive rtol = atol = 1e-4), inputs resident in HBM,
    CUDA-event timed, L2 flushed between iterations."""
    out = {}
    ffjord = ffjord_chain(m)
    cases = [
        # name, ICNF kwargs, precision, batch, mode
        ("config1_usage_B1024", dict(nvariables=1), "fp32", 1024, m.TestMode()),                        # examples/usage.jl shape, tiny
        ("config1_usage_B1024_vcabm", dict(nvariables=1), "fp32", 1024, m.TestMode()),                  # the same with the reference's default alg
        ("config2_moons_B65536", dict(nvariables=2, naugments=0), "fp32", 65536, m.TestMode()),         # tiny
        ("config2_moons_w64_B65536", dict(nvariables=2, naugments=0, n_hidden=64), "fp32", 65536, m.TestMode()),   # 3-64-64-2
        ("config2_moons_w64_B65536_hutch", dict(nvariables=2, naugments=0, n_hidden=64), "fp32", 65536, m.TrainMode(True)),   # same net, Hutchinson + regularisers
        ("config3_gmm16_B262144", dict(nvariables=16, naugments=0), "fp32", 262144, m.TestMode()),      # 17-68-68-16
        ("config3_gmm16_B262144_bf16x3tc", dict(nvariables=16, naugments=0), "bf16x3_tc", 262144, m.TestMode()),
        ("config5_cond64_B65536_fp32", dict(nvariables=64, naugments=0, nconditions=32), "fp32", 65536, m.TestMode()),   # 97-388-388-64
        ("config5_cond64_B65536_bf16tc", dict(nvariables=64, naugments=0, nconditions=32), "bf16_tc", 65536, m.TestMode()),
        ("config5_cond64_B65536_bf16x3tc", dict(nvariables=64, naugments=0, nconditions=32), "bf16x3_tc", 65536, m.TestMode()),
        ("config4_ffjord784_B8192_fp32", dict(nvariables=784, naugments=0, nn=ffjord), "fp32", 8192, m.TrainMode(False)),
        ("config4_ffjord784_B8192_bf16tc", dict(nvariables=784, naugments=0, nn=ffjord), "bf16_tc", 8192, m.TrainMode(False)),
        # split bf16 (hi + lo, 3 MMAs per K step): tensor cores at fp32-level accuracy, same step count as fp32
        ("config4_ffjord784_B8192_bf16x3tc", dict(nvariables=784, naugments=0, nn=ffjord), "bf16x3_tc", 8192, m.TrainMode(False)),
        # with a fixed step (8 steps, 48 RHS calls) bf16 rounding cannot inflate the step count
        ("config4_ffjord784_B8192_bf16tc_fixed8", dict(nvariables=784, naugments=0, nn=ffjord), "bf16_tc", 8192, m.TrainMode(False)),
    ]
    for name, kw, prec, B, mode in cases:
        sol = dict(adaptive=False, dt=0.125) if name.endswith("_fixed8") else (dict(alg="VCABM") if name.endswith("_vcabm") else {})
        icnf = _icnf(m, local, prec, **kw)
        rng = np.random.default_rng(7)
        theta, _ = m.setup(rng, icnf)
        err = parity_gate(m, icnf, mode, theta, "inference", sol)
        xs = torch.from_numpy(rng.standard_normal((B, icnf.nvariables)).astype(np.float32)).to(dev)
        args = (xs.t(),)
        if icnf.nconditions:
            args += (torch.from_numpy(rng.standard_normal((B, icnf.nconditions)).astype(np.float32)).to(dev).t(),)
        ms = _timed_calls(lambda: m.inference(icnf, mode, *args, theta, {}, seed=3, **sol), flush_buf)
        st = icnf.check_last()
        P = sum(icnf.sizes[i] * icnf.sizes[i + 1] for i in range(len(icnf.sizes) - 1))
        flop_rhs = 4 * P if not isinstance(mode, m.TestMode) else 2 * P + 2 * icnf.sizes[1] * icnf.sizes[2]
        out[name] = {"logp_evals_per_sec": B / (ms * 1e-3), "ms_per_call": ms, "kernel_family": icnf.solve_path(mode),
                     "mode": repr(mode), "solver_steps": st.naccept, "rhs_calls": st.nf, "parity_err_subbatch": err,
                     "algorithmic_tflops": flop_rhs * st.nf * B / (ms * 1e-3) / 1e12}
        del icnf
    return out


def config5_sweep(m, local, dev, flush_buf):
    """BASELINE.json configs[4]: CondICNF 64-D + 32-D conditioning, batched generate() and inference over batch sizes
    (97-388-388-64, TestMode = exact trace, adaptive Tsit5), split-bf16 tensor cores and the fp32 family."""
    out = {}
    for prec in ("bf16x3_tc", "fp32"):
        icnf = _icnf(m, local, prec, nvariables=64, naugments=0, nconditions=32)
        rng = np.random.default_rng(7)
        theta, _ = m.setup(rng, icnf)
        e_inf = parity_gate(m, icnf, m.TestMode(), theta, "inference", {})
        e_gen = parity_gate(m, icnf, m.TestMode(), theta, "generate", {})
        for B in (1024, 4096, 16384, 65536, 262144):
            if prec == "fp32" and B > 65536:
                continue
            xs = torch.from_numpy(rng.standard_normal((B, 64)).astype(np.float32)).to(dev)
            ys = torch.from_numpy(rng.standard_normal((B, 32)).astype(np.float32)).to(dev)
            ms_i = _timed_calls(lambda: m.inference(icnf, m.TestMode(), xs.t(), ys.t(), theta, {}), flush_buf, K=2, W=1)
            st_i = icnf.check_last()
            ms_g = _timed_calls(lambda: m.generate(icnf, m.TestMode(), ys.t(), theta, {}, B, seed=9), flush_buf, K=2, W=1)
            st_g = icnf.check_last()
            out[f"B{B}_{prec}"] = {"inference_evals_per_sec": B / (ms_i * 1e-3), "inference_ms": ms_i, "inference_steps": st_i.naccept,
                                   "generate_samples_per_sec": B / (ms_g * 1e-3), "generate_ms": ms_g, "generate_steps": st_g.naccept,
                                   "parity_err_subbatch": {"inference": e_inf, "generate": e_gen}, "kernel_family": icnf.kernel_family}
        del icnf
    return out


def w64_training(m, local, dev, flush_buf):
    """config 2's other width: 3-64-64-2 RNODE training step at batch 65 536 -- forward solve in one launch
    (narrow kernel, checkpoints), reverse sweep on the fp32 SGEMMs of the generic family."""
    icnf = _icnf(m, local, "fp32", nvariables=2, naugments=0, n_hidden=64, rng=4321)
    rng = np.random.default_rng(7)
    theta, _ = m.setup(rng, icnf)
    err = parity_gate(m, icnf, m.TrainMode(True), theta, "train", {}, nb=128)
    theta_d = torch.from_numpy(theta).to(dev)
    B = 65536
    xs = torch.from_numpy(np.ascontiguousarray(two_moons(B, seed=3).T.astype(np.float32))).to(dev)
    ms = _timed_calls(lambda: m.loss_and_gradient(icnf, m.TrainMode(True), xs.t(), theta_d, {}, seed=3), flush_buf)
    st = icnf.check_last()
    return {"train_samples_per_sec": B / (ms * 1e-3), "ms_per_step": ms,
            "kernel_family": icnf.solve_path(m.TrainMode(True)) + " forward solve, " + icnf.kernel_family + " reverse sweep",
            "mode": "TrainMode(True) loss + gradient", "solver_steps": st.naccept, "rhs_calls": st.nf, "solver_status": st.status,
            "parity_err_subbatch": err}


def config4_training(m, local, dev, flush_buf, rank, world, precisions):
    """BASELINE.json configs[3]: 784-D FFJORD (785-512-512-512-784 softplus) RNODE TRAINING step, 8192 samples per GPU,
    adaptive Tsit5, loss + gradient of all 1.33 M parameters + (N > 1) one all-reduce of the 5.3 MB [gradient; loss]."""
    import torch.distributed as dist
    out = {}
    B = 8192
    P = 785 * 512 + 512 * 512 * 2 + 512 * 784
    for prec in precisions:
        icnf = _icnf(m, local, prec, nvariables=784, naugments=0, nn=ffjord_chain(m), rng=4321)
        rng = np.random.default_rng(7)
        theta, _ = m.setup(rng, icnf)
        err = parity_gate(m, icnf, m.TrainMode(True), theta, "train", {}, nb=128) if rank == 0 else None
        theta_d = torch.from_numpy(theta).to(dev)
        xs = torch.from_numpy(np.random.default_rng(100 + rank).uniform(0, 1, (B, 784)).astype(np.float32)).to(dev)   # U[0,1) "pixels"

        def step():
            return m.dp_loss_and_gradient(icnf, m.TrainMode(True), xs.t(), theta_d, {}, rank=rank, world=world,
                                          global_batch=B * world, seed=3)
        for _ in range(2):
            step()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        K, ms = 3, 0.0
        for _ in range(K):
            flush_buf.zero_()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            l, g = step()
            b.record()
            torch.cuda.synchronize()
            ms += a.elapsed_time(b)
        t = torch.tensor([ms / K], device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
        st = icnf.check_last()
        assert torch.isfinite(g).all() and math.isfinite(float(l))
        flop = 72.0 * P * B * st.naccept + 4.0 * P * B * st.nf      # reverse sweep + forward solve, algorithmic
        out[f"config4_ffjord784_B8192_train_{prec}"] = {
            "train_samples_per_sec": B * world / (ms * 1e-3), "ms_per_step": ms, "kernel_family": icnf.kernel_family,
            "mode": "TrainMode(True) loss + gradient" + (f" + all-reduce of {4 * (icnf.n_params + 1)} bytes over {world} GPUs" if world > 1 else ""),
            "solver_steps": st.naccept, "rhs_calls": st.nf, "solver_status": st.status, "parity_err_subbatch": err,
            "algorithmic_tflops_per_gpu": flop / (ms * 1e-3) / 1e12, "n_gpus": world}
        del icnf
    return out


def device_adam_variant(m, local, dev, flush_buf, B):
    """the headline step holds theta constant.  This variant changes theta EVERY step: device-resident
    parameters, loss + gradient, then the Adam + weight-decay update on the device (icnf_adam_step_dev), so the
    parameter refresh of the next step (icnf_set_params_dev) is inside the timed region."""
    icnf = m.ICNF(nvariables=2, naugments=0, device=local, epsdist="rademacher", rng=1234)
    theta = torch.from_numpy(init_theta()).to(dev)
    mom, var = torch.zeros_like(theta), torch.zeros_like(theta)
    xs = torch.from_numpy(np.ascontiguousarray(two_moons(B, seed=1).T)).to(dev)
    step_no = [0]

    def step():
        step_no[0] += 1
        l, g = m.loss_and_gradient(icnf, m.TrainMode(True), xs.t(), theta, {}, seed=1000 + step_no[0])
        rc = m.lib.icnf_adam_step_dev(theta.data_ptr(), g.data_ptr(), mom.data_ptr(), var.data_ptr(), theta.numel(), step_no[0],
                                      1e-3, 0.9, 0.999, 1e-8, 1e-4, torch.cuda.current_stream().cuda_stream)
        assert rc == 0
    ms = _timed_calls(step, flush_buf, K=10, W=5)
    icnf.check_last()
    return {"train_samples_per_sec": B / (ms * 1e-3), "ms_per_step": ms,
            "note": "theta updated on the device every step (Adam + WeightDecay), re-read by the library each step"}


# ------------------------------------------------------------------ our arm
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=BATCH_PER_GPU, help="samples per GPU per step")
    ap.add_argument("--cpu-batch", type=int, default=65536, help="batch of the bounded CPU sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the side measurements (other BASELINE configs)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        return run_reference(args)

    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product has no CPU path")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    import cnf_b200 as m
    B = args.batch
    # the STEER stream is seeded identically on every rank: the unsharded solve draws ONE t1 for the whole batch
    # (base_icnf.jl:23-43), so the shards of a data-parallel step must integrate to the same t1
    icnf = m.ICNF(nvariables=2, naugments=0, device=local, epsdist="rademacher", rng=1234)
    assert icnf.kernel_family == "tiny"
    if world > 1:
        m.group_join(icnf, rank, world)       # communicator inside the library (NCCL / NVLink peer memory)
    theta = init_theta()
    xs_all = two_moons(B * world, seed=1)
    xs_host = torch.from_numpy(np.ascontiguousarray(xs_all[:, rank * B:(rank + 1) * B].T)).pin_memory()   # (B, 2) = 2 x B column-major
    xs_dev = xs_host.to(dev)
    mode = m.TrainMode(True)
    flush_buf = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)     # > 126 MB L2

    def step_resident(i):
        # local loss/gradient on this rank's columns, then ONE all-reduce of [gradient; loss] when N > 1
        return m.dp_loss_and_gradient(icnf, mode, xs_dev.t(), theta, {}, rank=rank, world=world, global_batch=B * world,
                                      seed=1000 + i)

    xs_np = xs_host.numpy().T       # 2 x B view of the pinned buffer, column-major

    def step_e2e(i):
        if world == 1:
            # the reference-facing call: host pointers in, host results out
            l, g = m.loss_and_gradient(icnf, mode, xs_np, theta, {}, seed=1000 + i, global_batch=B)
            return float(l), g
        xd = xs_host.to(dev, non_blocking=True)
        l, g = m.dp_loss_and_gradient(icnf, mode, xd.t(), theta, {}, rank=rank, world=world, global_batch=B * world,
                                      seed=1000 + i)
        both = icnf._grad_loss_buf.cpu().numpy()       # [gradient; loss] share one device buffer: one device-to-host copy
        return float(both[-1]), both[:-1]

    def timed(fn, K, W):
        for i in range(W):
            fn(i)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
        launches0 = icnf.launch_count
        for i in range(K):
            flush_buf.zero_()                       # L2 flush between timed iterations (not timed)
            evs[i][0].record()
            fn(W + i)
            evs[i][1].record()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        ms = sum(a.elapsed_time(b) for a, b in evs)
        t = torch.tensor([ms], device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item()), icnf.launch_count - launches0

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ms_total, launches = timed(step_resident, args.steps, args.warmup)
    clocks = sampler.stop() if rank == 0 else None
    icnf.check_last()          # the last timed solve reached t1 (a failed solve would have skipped its backward sweep)

    # e2e: host-side wall clock around the reference-facing call (it returns when results are on the host)
    def timed_wall(fn, K, W):
        for i in range(W):
            fn(i)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        for i in range(K):
            fn(W + i)
        torch.cuda.synchronize()
        el = time.perf_counter() - t0
        t = torch.tensor([el], device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())
    e2e_s = timed_wall(step_e2e, args.steps, args.warmup)

    # per-kernel device times (CUDA events inside the library, on the launching stream)
    icnf.set_profiling(True)
    kt = {"forward": [], "backward": [], "loss_sum": [], "grad_reduce": []}
    nf_list, nacc_list = [], []
    for i in range(args.steps):
        flush_buf.zero_()
        l, g = m.loss_and_gradient(icnf, mode, xs_np, theta, {}, seed=1000 + args.warmup + i, sample_offset=rank * B,
                                   global_batch=B * world)
        for k, v in icnf.kernel_times_ms().items():
            kt[k].append(v)
        nf_list.append(icnf.last_stats.nf)
        nacc_list.append(icnf.last_stats.naccept)
    icnf.set_profiling(False)

    extras = {}
    if not args.no_extras:
        # BASELINE configs[3]: config-4 training at THIS N (every rank takes part: the gradient all-reduce is 5.3 MB)
        extras.update(config4_training(m, local, dev, flush_buf, rank, world, ("bf16x3_tc", "fp32") if world == 1 else ("bf16x3_tc",)))

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "MEASURED_PEAKS.json hbm_gbs (burst copy), of measured" if peaks else "fallback 6650 GB/s, of fallback"
    fp32_peak = m.measure_fp32_peak(local)

    ms_step = ms_total / args.steps
    value = B * world / (ms_step * 1e-3)
    bwd_ms = float(np.mean(kt["backward"]))
    fwd_ms = float(np.mean(kt["forward"]))
    nacc = float(np.mean(nacc_list))
    nf = float(np.mean(nf_list))
    # dominant kernel = backward.  Algorithmic bytes per sample per launch: eps regenerated in-kernel,
    # weights in the parameter bank; it reads the stage-input checkpoints (6 records of D' floats per
    # accepted step, plus the final state).
    bwd_bytes = B * (6 * nacc + 1) * 2 * 4.0
    bwd_flop = B * nacc * 6 * (8 * PW + 4 * PW)                     # 6 stages x 12 P flop per step per sample (stage inputs are checkpointed)
    fwd_flop = B * nf * 4 * PW                                        # Hutchinson RHS = 4 P flop
    achieved_gbs = bwd_bytes / (bwd_ms * 1e-3) / 1e9
    traffic = None
    try:   # DRAM bytes per launch of the dominant kernel from the committed ncu --set full capture
        import glob
        tr = json.load(open(sorted(glob.glob(os.path.join(ROOT, "profiles", "r*_traffic.json")))[-1]))["tiny::backward_sp_kernel"]
        traffic = tr["dram_bytes_per_launch"] * (B / tr["batch"])
    except Exception:
        pass
    bwd_tf = bwd_flop / (bwd_ms * 1e-3) / 1e12
    fwd_tf = fwd_flop / (fwd_ms * 1e-3) / 1e12
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic", "config": make_config(B, world),
        "timing": "per-step CUDA events on the launching stream, L2 flushed (256 MiB write) between timed iterations, max over ranks",
        "solver": {"steps_mean": nacc, "rhs_calls_per_solve_mean": nf},
        "e2e": {"value": B * world * args.steps / e2e_s, "unit": UNIT, "h2d_bytes_per_step": B * 2 * 4,
                "d2h_bytes_per_step": 4 * (NPARAMS + 1) + 24,
                "path": "icnf_loss_grad (host-pointer C ABI): pinned host xs in, loss+gradient+stats out" if world == 1
                else "pinned H2D copy + icnf_loss_grad_dp_dev (all-reduce inside the library) + D2H of loss and gradient"},
        "gpu_launches": launches,
        "clocks": clocks,
        # the bound that binds: the narrow-MLP kernels are FP32-issue bound (AI 20 flop/B against a
        # machine balance of 11), so the primary roofline is the FP32 FMA pipe; the HBM view is kept beside it
        "roofline": {"bound": "fp32_fma", "kernel": "tiny::backward_sp_kernel", "achieved": bwd_tf, "peak": fp32_peak,
                     "unit": "TFLOP/s", "frac": bwd_tf / fp32_peak, "traffic": traffic,
                     "peak_source": "FFMA-chain microbenchmark in this run (icnf_measure_fp32_peak); MEASURED_PEAKS.json has no FP32 figure",
                     "forward_kernel": {"kernel": "tiny::solve_adaptive_kernel", "achieved": fwd_tf, "frac": fwd_tf / fp32_peak}},
        "roofline_hbm": {"bound": "hbm", "kernel": "tiny::backward_sp_kernel", "achieved": achieved_gbs, "peak": hbm_peak, "unit": "GB/s",
                         "frac": achieved_gbs / hbm_peak, "traffic": traffic, "peak_source": peak_src},
        "kernel_ms": {k: float(np.mean(v)) for k, v in kt.items()},
    }
    if not args.no_extras:
        if world == 1:
            extras.update(logp_extras(m, local, dev, flush_buf))
            extras["config5_generate_inference_sweep"] = config5_sweep(m, local, dev, flush_buf)
            extras["config2_train_device_adam"] = device_adam_variant(m, local, dev, flush_buf, B)
            extras["config2_moons_w64_B65536_train"] = w64_training(m, local, dev, flush_buf)
            c4 = extras.get("config4_ffjord784_B8192_train_bf16x3_tc")
            if c4 and peaks.get("bf16_tflops_sustained"):
                # tensor-pipe roofline of the wide path: algorithmic flop (x3 physical MMAs in split precision are
                # NOT counted) against the measured sustained bf16 GEMM rate
                line["roofline_tensor"] = {"bound": "tensor", "kernel": "tc::tc_gemm_kernel (config-4 training step)",
                                           "achieved": c4["algorithmic_tflops_per_gpu"], "peak": peaks["bf16_tflops_sustained"],
                                           "unit": "TFLOP/s", "frac": c4["algorithmic_tflops_per_gpu"] / peaks["bf16_tflops_sustained"],
                                           "peak_source": "MEASURED_PEAKS.json bf16_tflops_sustained, of measured"}
            c3 = extras.get("config3_gmm16_B262144")
            if c3:
                # FP32 roofline of the single-launch exact-trace solve (csrc/narrow_kernel.cuh): algorithmic flop per RHS and sample
                # = 2 P + 2 n1 n2 (closed-form trace; the reference's D' one-hot pullbacks would count 2 P (1 + D'))
                line["roofline_fp32_config3"] = {"bound": "fp32_fma", "kernel": "narrow::solve_kernel (config 3, whole adaptive solve)",
                                                 "achieved": c3["algorithmic_tflops"], "peak": fp32_peak, "unit": "TFLOP/s",
                                                 "frac": c3["algorithmic_tflops"] / fp32_peak,
                                                 "peak_source": "FFMA-chain microbenchmark in this run (icnf_measure_fp32_peak)"}
        line["extras"] = extras
    if not args.no_cpu_baseline and world == 1:
        threads = os.cpu_count() or 1
        rate, ms, nfc = cpu_loss_grad_rate(args.cpu_batch, 5, 1, threads)
        line["cpu_baseline"] = {"value": rate, "unit": UNIT, "cores": threads, "kind": "port",
                                "sample": f"5 training steps of a {args.cpu_batch}-sample batch of the same workload on the box's host cores (oracle/: torch CPU fp32 + autograd; nf={nfc})"}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
