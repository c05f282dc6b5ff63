#!/usr/bin/env python
"""bench.py -- headline measurement of the B200 hot path.

Workload (BASELINE.json configs[1]): 2-D two-moons RNODE, nvars 2, naug 0,
3 -> 12 -> 12 -> 2 softplus MLP (the reference's default width 4 n_in), one
Hutchinson Rademacher probe drawn in-kernel, regularisers lambda1 = lambda2 = 0.01,
STEER end time, Tsit5 adaptive at reltol = abstol = 1e-4 (the reference's default
tolerances), batch 65 536 per GPU.  A "step" is one training step: loss + gradient
w.r.t. all parameters (+ one NCCL all-reduce of the gradient when N > 1).

    python bench.py --gpus N --steps K --warmup W            # our arm
    python bench.py --impl reference --gpus N --steps K ...  # CPU restatement arm

Prints ONE JSON line (rank 0).  `value` = samples/s with the batch resident in
HBM; `e2e` = the same step through the host-pointer C ABI (pinned host xs in,
loss + gradient out) -- see DESIGN.md "Measurement".
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
    line = {
        "impl": "reference", "metric": METRIC, "value": rate, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "batch_per_step": B, "note": "CPU restatement of the reference path (oracle/, torch CPU fp32 + autograd); the Julia package cannot run here"},
        "cpu_baseline": {"value": rate, "unit": UNIT, "cores": threads, "kind": "port",
                         "sample": f"{args.steps} training steps of a {B}-sample batch of the same workload (nf={nf} RHS calls per solve)"},
        "e2e": {"value": rate, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------ log p(x) side measurements
def logp_extras(m, local, dev, flush_buf):
    """The other half of BASELINE.json's metric: log-density evaluations per second
    (TestMode = exact trace, Tsit5 adaptive rtol = atol = 1e-4), inputs resident in HBM,
    CUDA-event timed, L2 flushed between iterations."""
    out = {}
    ffjord = m.Chain(m.Dense(785, 512, "softplus"), m.Dense(512, 512, "softplus"), m.Dense(512, 512, "softplus"), m.Dense(512, 784))
    cases = [
        # name, ICNF kwargs, batch, mode
        ("config1_usage_B1024", dict(nvariables=1), 1024, m.TestMode()),                     # examples/usage.jl shape, tiny
        ("config2_moons_B65536", dict(nvariables=2, naugments=0), 65536, m.TestMode()),      # tiny
        ("config3_gmm16_B262144", dict(nvariables=16, naugments=0), 262144, m.TestMode()),   # 17-68-68-16, generic fp32
        ("config3_gmm16_B262144_bf16tc", dict(nvariables=16, naugments=0, precision="bf16_tc"), 262144, m.TestMode()),
        ("config3_gmm16_B262144_bf16x3tc", dict(nvariables=16, naugments=0, precision="bf16x3_tc"), 262144, m.TestMode()),
        ("config5_cond64_B65536_fp32", dict(nvariables=64, naugments=0, nconditions=32), 65536, m.TestMode()),   # 97-388-388-64
        ("config5_cond64_B65536_bf16tc", dict(nvariables=64, naugments=0, nconditions=32, precision="bf16_tc"), 65536, m.TestMode()),
        ("config5_cond64_B65536_bf16x3tc", dict(nvariables=64, naugments=0, nconditions=32, precision="bf16x3_tc"), 65536, m.TestMode()),
        ("config4_ffjord784_B8192_fp32", dict(nvariables=784, naugments=0, nn=ffjord), 8192, m.TrainMode(False)),
        ("config4_ffjord784_B8192_bf16tc", dict(nvariables=784, naugments=0, nn=ffjord, precision="bf16_tc"), 8192, m.TrainMode(False)),
        # split bf16 (hi + lo, 3 MMAs per K step): tensor cores at fp32-level accuracy, same step count as fp32
        ("config4_ffjord784_B8192_bf16x3tc", dict(nvariables=784, naugments=0, nn=ffjord, precision="bf16x3_tc"), 8192, m.TrainMode(False)),
        # the same two with a fixed step (8 steps, 48 RHS calls): bf16 rounding cannot inflate the step count
        ("config4_ffjord784_B8192_fp32_fixed8", dict(nvariables=784, naugments=0, nn=ffjord), 8192, m.TrainMode(False)),
        ("config4_ffjord784_B8192_bf16tc_fixed8", dict(nvariables=784, naugments=0, nn=ffjord, precision="bf16_tc"), 8192, m.TrainMode(False)),
    ]
    for name, kw, B, mode in cases:
        sol = dict(adaptive=False, dt=0.125) if name.endswith("_fixed8") else {}
        icnf = m.ICNF(device=local, epsdist="rademacher", **kw)
        rng = np.random.default_rng(7)
        theta, _ = m.setup(rng, icnf)
        xs = torch.from_numpy(rng.standard_normal((B, icnf.nvariables)).astype(np.float32)).to(dev)
        args = (xs.t(),)
        if icnf.nconditions:
            args += (torch.from_numpy(rng.standard_normal((B, icnf.nconditions)).astype(np.float32)).to(dev).t(),)
        for _ in range(2):
            m.inference(icnf, mode, *args, theta, {}, seed=3, **sol)
        torch.cuda.synchronize()
        K, ms = 3, 0.0
        for _ in range(K):
            flush_buf.zero_()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            m.inference(icnf, mode, *args, theta, {}, seed=3, **sol)
            b.record()
            torch.cuda.synchronize()
            ms += a.elapsed_time(b)
        st = icnf.last_stats
        P = sum(icnf.sizes[i] * icnf.sizes[i + 1] for i in range(len(icnf.sizes) - 1))
        flop_rhs = 4 * P if not isinstance(mode, m.TestMode) else 2 * P + 2 * icnf.sizes[1] * icnf.sizes[2]
        out[name] = {"logp_evals_per_sec": B * K / (ms * 1e-3), "ms_per_call": ms / K, "kernel_family": icnf.kernel_family,
                     "mode": repr(mode), "solver_steps": st.naccept, "rhs_calls": st.nf,
                     "algorithmic_tflops": flop_rhs * st.nf * B / (ms / K * 1e-3) / 1e12}
        del icnf
    # BASELINE config 4 as a TRAINING step (loss + gradient, RNODE regularisers, adaptive Tsit5): fp32 family, and the
    # split-precision tensor-core forward with the fp32 reverse sweep
    for name, prec in (("config4_ffjord784_B8192_train_fp32", "fp32"), ("config4_ffjord784_B8192_train_bf16x3tc", "bf16x3_tc")):
        B = 8192
        icnf = m.ICNF(device=local, epsdist="rademacher", nvariables=784, naugments=0, nn=ffjord, precision=prec)
        rng = np.random.default_rng(7)
        theta, _ = m.setup(rng, icnf)
        theta_d = torch.from_numpy(theta).to(dev)
        xs = torch.from_numpy(rng.standard_normal((B, 784)).astype(np.float32)).to(dev)
        for _ in range(2):
            m.loss_and_gradient(icnf, m.TrainMode(True), xs.t(), theta_d, {}, seed=3)
        torch.cuda.synchronize()
        K, ms = 3, 0.0
        for _ in range(K):
            flush_buf.zero_()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            m.loss_and_gradient(icnf, m.TrainMode(True), xs.t(), theta_d, {}, seed=3)
            b.record()
            torch.cuda.synchronize()
            ms += a.elapsed_time(b)
        st = icnf.last_stats
        out[name] = {"train_samples_per_sec": B * K / (ms * 1e-3), "ms_per_call": ms / K, "kernel_family": icnf.kernel_family,
                     "mode": "TrainMode(True) loss + gradient", "solver_steps": st.naccept, "rhs_calls": st.nf,
                     "logp_evals_per_sec": 0.0, "algorithmic_tflops": 0.0}
        del icnf
    return out


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
    ap.add_argument("--no-extras", action="store_true", help="skip the log p(x) evals/s side measurements")
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
    theta = init_theta()
    xs_all = two_moons(B * world, seed=1)
    xs_host = torch.from_numpy(np.ascontiguousarray(xs_all[:, rank * B:(rank + 1) * B].T)).pin_memory()   # (B, 2) = 2 x B column-major
    xs_dev = xs_host.to(dev)
    mode = m.TrainMode(True)
    flush_buf = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)     # > 126 MB L2

    def step_resident(i):
        # local loss/gradient on this rank's columns, then ONE all-reduce of [gradient; loss] (NCCL) when N > 1
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
        return float(l.cpu()), g.cpu().numpy()

    def timed(fn, K, W):
        for i in range(W):
            fn(i)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
        launches0 = icnf.launch_count
        wall0 = time.perf_counter()
        for i in range(K):
            flush_buf.zero_()                       # L2 flush between timed iterations (not timed)
            evs[i][0].record()
            fn(W + i)
            evs[i][1].record()
        torch.cuda.synchronize()
        wall = time.perf_counter() - wall0
        if world > 1:
            dist.barrier()
        ms = sum(a.elapsed_time(b) for a, b in evs)
        t = torch.tensor([ms], device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item()), icnf.launch_count - launches0, wall

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ms_total, launches, _ = timed(step_resident, args.steps, args.warmup)
    clocks = sampler.stop() if rank == 0 else None
    stats = icnf.last_stats if world == 1 else None
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
        tr = json.load(open(os.path.join(ROOT, "profiles", "r01_traffic.json")))["tiny::backward_sp_kernel"]
        traffic = tr["dram_bytes_per_launch"] * (B / tr["batch"])
    except Exception:
        pass
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": {"workload": WORKLOAD, "batch_per_gpu": B, "global_batch": B * world, "parallelism": f"dp{world}",
                   "timing": "per-step CUDA events on the launching stream, L2 flushed (256 MiB write) between timed iterations, max over ranks",
                   "solver_steps_mean": nacc, "rhs_calls_per_solve_mean": nf},
        "e2e": {"value": B * world * args.steps / e2e_s, "unit": UNIT, "h2d_bytes_per_step": B * 2 * 4,
                "d2h_bytes_per_step": 4 * (NPARAMS + 1) + 24,
                "path": "icnf_loss_grad (host-pointer C ABI): pinned host xs in, loss+gradient+stats out" if world == 1
                else "pinned H2D copy + icnf_loss_grad_dev + NCCL all-reduce + D2H of loss and gradient"},
        "gpu_launches": launches,
        "clocks": clocks,
        "roofline": {"bound": "hbm", "kernel": "tiny::backward_sp_kernel", "achieved": achieved_gbs, "peak": hbm_peak, "unit": "GB/s",
                     "frac": achieved_gbs / hbm_peak, "traffic": traffic, "peak_source": peak_src,
                     "note": "narrow-MLP path is FP32-issue bound, not HBM bound (SURVEY 8(d)); see roofline_fp32"},
        "roofline_fp32": {"bound": "fp32_fma", "kernel": "tiny::backward_sp_kernel", "achieved": bwd_flop / (bwd_ms * 1e-3) / 1e12,
                          "peak": fp32_peak, "unit": "TFLOP/s", "frac": bwd_flop / (bwd_ms * 1e-3) / 1e12 / fp32_peak,
                          "peak_source": "FFMA-chain microbenchmark in this run (icnf_measure_fp32_peak)",
                          "forward_kernel": {"kernel": "tiny::solve_adaptive_kernel", "achieved": fwd_flop / (fwd_ms * 1e-3) / 1e12,
                                             "frac": fwd_flop / (fwd_ms * 1e-3) / 1e12 / fp32_peak}},
        "kernel_ms": {k: float(np.mean(v)) for k, v in kt.items()},
    }
    if world == 1 and not args.no_extras:
        line["extras"] = logp_extras(m, local, dev, flush_buf)
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
