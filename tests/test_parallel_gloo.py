"""world_size-2 gloo test of the multi-GPU protocol: shards sum to
the whole.  The local compute is the CPU oracle here (no GPU in this container);
on the GPU box the same helper runs the CUDA path under NCCL (bench.py --gpus N)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _oracle_local(icnf, mode, xs, theta, st, sample_offset=0, global_batch=0, seed=0, **kw):
    """CPU stand-in with the C ABI's sharding contract: noise by global column index,
    mean over the GLOBAL batch."""
    from oracle import icnf_oracle as O
    from oracle import philox as P
    om = icnf
    n = xs.shape[1]
    eps = torch.tensor(P.rademacher(seed, om.d, n, offset=sample_offset), dtype=torch.float64)
    val, g, _ = O.loss_grad(om, O.TRAIN_REG, torch.tensor(xs, dtype=torch.float64), torch.tensor(theta, dtype=torch.float64),
                            eps, opts=O.SolverOpts(adaptive=False, dt=0.25))
    scale = n / float(global_batch)
    return float(val) * scale, (g * scale).numpy().astype(np.float32)


def _worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import cnf_b200 as m
    from oracle import icnf_oracle as O
    om = O.OracleICNF(nvars=2, naug=0)
    theta = O.init_params(om, 1, np.float32, bias_scale=0.2)
    B = 37
    xs = np.random.default_rng(0).standard_normal((2, B)).astype(np.float32)
    lo, hi = m.shard_bounds(B, rank, world)
    l, g = m.dp_loss_and_gradient(om, None, xs[:, lo:hi], theta, {}, rank=rank, world=world, global_batch=B,
                                  local_fn=_oracle_local, seed=11)
    if rank == 0:
        lw, gw = _oracle_local(om, None, xs, theta, {}, sample_offset=0, global_batch=B, seed=11)
        out.put((l, g, lw, gw))
    dist.barrier()
    dist.destroy_process_group()


def test_shard_bounds_cover_the_batch():
    sys.path.insert(0, ROOT)
    import cnf_b200 as m
    for n in (0, 1, 7, 64, 65537):
        for world in (1, 2, 3, 8):
            parts = [m.shard_bounds(n, r, world) for r in range(world)]
            assert parts[0][0] == 0 and parts[-1][1] == n
            assert all(parts[i][1] == parts[i + 1][0] for i in range(world - 1))
            sizes = [hi - lo for lo, hi in parts]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        m.shard_bounds(4, 2, 2)


def test_two_rank_gradient_allreduce_matches_unsharded():
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    l, g, lw, gw = out.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert l == pytest.approx(lw, rel=1e-5)
    np.testing.assert_allclose(g, gw, rtol=1e-4, atol=1e-6)
