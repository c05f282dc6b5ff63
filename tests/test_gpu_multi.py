"""Multi-GPU tests of the library's own communicator (icnf_group_join / icnf_create_group / icnf_loss_grad_dp):
need at least two GPUs on the box (`gpurun --gpus 2`), skipped otherwise.  The shards' data-parallel gradient must
equal the unsharded single-GPU gradient (fixed-step solves: exactly the same discrete object), on both exchange
paths -- NVLink peer memory inside the gradient-reduction kernel (narrow MLPs) and ncclAllReduce (wide ones)."""
import os
import socket
import sys
import threading

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

needs2 = pytest.mark.skipif(not torch.cuda.is_available() or torch.cuda.device_count() < 2, reason="needs two GPUs")


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _case(m, which, device):
    if which == "tiny":
        return m.ICNF(nvariables=2, naugments=0, device=device), 1001
    if which == "narrow":       # ICNF(nvariables = 3): 8-32-32-7, not in the tiny whitelist -> generic family, single-launch solves
        return m.ICNF(nvariables=3, device=device), 1001
    nn = m.Chain(m.Dense(97, 128, "softplus"), m.Dense(128, 160, "softplus"), m.Dense(160, 128, "softplus"), m.Dense(128, 96))
    return m.ICNF(nvariables=96, naugments=0, nn=nn, device=device, precision="bf16x3_tc"), 515


def _inputs(icnf, B):
    from tests.helpers import make_inputs
    return make_inputs(icnf, B)


def _worker(rank, world, port, which, out):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    import torch.distributed as dist
    dist.init_process_group("gloo", rank=rank, world_size=world)     # only ships the 128-byte group id
    torch.cuda.set_device(rank)
    import cnf_b200 as m
    icnf, B = _case(m, which, rank)
    om, theta, xs, eps, _ = _inputs(icnf, B)
    m.group_join(icnf, rank, world)
    info = m.group_info(icnf)
    lo, hi = m.shard_bounds(B, rank, world)
    sol = dict(adaptive=False, dt=0.25)
    dev = f"cuda:{rank}"
    xd = torch.tensor(np.ascontiguousarray(xs[:, lo:hi])).to(dev)
    ed = torch.tensor(np.ascontiguousarray(eps[:, lo:hi])).to(dev)
    res = []
    for rep in range(3):     # several steps: the exchange buffers are reused with alternating parity
        l, g = m.dp_loss_and_gradient(icnf, m.TrainMode(True), xd, theta, {}, rank=rank, world=world, global_batch=B,
                                      eps=ed, tspan=icnf.tspan, **sol)
        res.append((float(l), g.cpu().numpy()))
    # host-pointer entry point of the same step
    lh, gh = m.loss_and_gradient(icnf, m.TrainMode(True), xs[:, lo:hi], theta, {}, eps=eps[:, lo:hi], tspan=icnf.tspan,
                                 sample_offset=lo, global_batch=B, data_parallel=True, **sol)
    exact = None
    if which in ("tiny", "narrow"):
        # exact mode: ADAPTIVE solve with the error norm of the global batch -> the unsharded solve's own steps
        m.group_set_global_norm(icnf, True)
        la, ga = m.dp_loss_and_gradient(icnf, m.TrainMode(True), xd, theta, {}, rank=rank, world=world, global_batch=B,
                                        eps=ed, tspan=icnf.tspan)
        sa = icnf.check_last()
        m.group_set_global_norm(icnf, False)
        lb, gb = m.dp_loss_and_gradient(icnf, m.TrainMode(True), xd, theta, {}, rank=rank, world=world, global_batch=B,
                                        eps=ed, tspan=icnf.tspan)
        exact = (float(la), ga.cpu().numpy(), sa.naccept, sa.nreject, sa.t_final, float(lb), gb.cpu().numpy())
    if rank == 0:
        solo, _ = _case(m, which, 0)
        lw, gw = m.loss_and_gradient(solo, m.TrainMode(True), xs, theta, {}, eps=eps, tspan=solo.tspan, **sol)
        if exact is not None:
            lwa, gwa = m.loss_and_gradient(solo, m.TrainMode(True), xs, theta, {}, eps=eps, tspan=solo.tspan)
            sw = solo.last_stats
            exact = exact + (float(lwa), gwa, sw.naccept, sw.nreject)
        out.put((res, (float(lh), gh), (float(lw), gw), info, exact))
    dist.barrier()
    dist.destroy_process_group()


@needs2
@pytest.mark.parametrize("which", ["tiny", "narrow", "wide_tc"])
def test_two_process_group_gradient_equals_unsharded(which):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, which, out)) for r in range(2)]
    for p in procs:
        p.start()
    res, host, whole, info, exact = out.get(timeout=300)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert info["n_ranks"] == 2
    if which == "tiny":
        assert info["peer_memory"], "NVLink peer-memory exchange should be available between two GPUs of one box"
    lw, gw = whole
    tol = 2e-6 if which == "tiny" else 2e-5       # summation order differs between 1 and 2 shards, nothing else
    if which == "narrow":
        assert info["peer_memory"]
    for l, g in res + [host]:
        assert abs(l - lw) <= 1e-5 * abs(lw)
        assert np.linalg.norm(g - gw) / np.linalg.norm(gw) < tol
    # the steps are identical, bit for bit (deterministic exchange)
    if which == "narrow":    # the fp32 weight-gradient SGEMMs of the generic family accumulate their split-K slices with atomics
        assert all(np.linalg.norm(res[0][1] - r[1]) <= 2e-6 * np.linalg.norm(res[0][1]) for r in res[1:])
    else:
        assert all(np.array_equal(res[0][1], r[1]) for r in res[1:])
    if exact is not None:
        la, ga, nacc, nrej, tf, lb, gb, lwa, gwa, nacc_w, nrej_w = exact
        # global error norm: same accepted / rejected steps as the unsharded adaptive solve, same gradient to rounding
        assert (nacc, nrej) == (nacc_w, nrej_w)
        assert abs(la - lwa) <= 1e-5 * abs(lwa) and np.linalg.norm(ga - gwa) / np.linalg.norm(gwa) < (5e-6 if which == "tiny" else 2e-5)
        # shard-local norm (default): agreement to solver tolerance only
        assert abs(lb - lwa) <= 1e-3 * abs(lwa) and np.linalg.norm(gb - gwa) / np.linalg.norm(gwa) < 1e-2


@needs2
def test_single_process_group_two_devices():
    """icnf_create_group: one process drives both GPUs (the natural Julia usage), one thread per device"""
    import cnf_b200 as m
    icnfs = [m.ICNF(nvariables=2, naugments=0, device=d) for d in range(2)]
    m.create_group(icnfs)
    B = 777
    om, theta, xs, eps, _ = _inputs(icnfs[0], B)
    sol = dict(adaptive=False, dt=0.25)
    outs = [None, None]

    def run(r):
        torch.cuda.set_device(r)
        lo, hi = m.shard_bounds(B, r, 2)
        outs[r] = m.loss_and_gradient(icnfs[r], m.TrainMode(True), xs[:, lo:hi], theta, {}, eps=eps[:, lo:hi], tspan=(0.0, 1.0),
                                      sample_offset=lo, global_batch=B, data_parallel=True, **sol)
    ths = [threading.Thread(target=run, args=(r,)) for r in range(2)]
    for t in ths:
        t.start()
    for t in ths:
        t.join(timeout=120)
    solo = m.ICNF(nvariables=2, naugments=0, device=1)        # a handle on the SECOND device of the process
    lw, gw = m.loss_and_gradient(solo, m.TrainMode(True), xs, theta, {}, eps=eps, tspan=(0.0, 1.0), **sol)
    for l, g in outs:
        assert abs(l - lw) <= 1e-5 * abs(lw)
        assert np.linalg.norm(g - gw) / np.linalg.norm(gw) < 2e-6
    assert np.array_equal(outs[0][1], outs[1][1])


def test_handles_on_two_devices_in_one_process_do_not_share_launch_state():
    """per-process statics (SM count, shared-memory attributes) must be per device"""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    import cnf_b200 as m
    from oracle import icnf_oracle as O
    from tests.helpers import norm_rel_err, t64
    for kw in (dict(nvariables=16, naugments=0), dict(nvariables=64, naugments=0, nconditions=32, n_hidden=256, precision="bf16x3_tc")):
        for dev in (0, 1):
            icnf = m.ICNF(device=dev, **kw)
            om, theta, xs, eps, ys = _inputs(icnf, 300)
            theta = (0.5 * theta).astype(np.float32)
            a = (xs, ys) if ys is not None else (xs,)
            lp, _ = m.inference(icnf, m.TestMode(), *a, theta, {})
            ref, _ = O.inference(om, O.TEST, t64(xs), t64(theta), None, t64(ys))
            assert norm_rel_err(lp, ref.numpy()) < 1e-4
