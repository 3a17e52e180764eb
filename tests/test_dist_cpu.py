"""Host-side data-parallel logic on CPU: world size 2 over gloo.  The gradient exchange must turn
per-shard (mean-of-shard) gradients into the full-batch gradient, shard_rows must partition the
batch, and log rows must average — checked with the oracle supplying real gradients."""
import os
import socket
import sys

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


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from cdgvae_b200 import dist as D
        from oracle import cdgvae_oracle as orc
        torch.set_num_threads(1)
        cfg = dict(dataset="adult", node=3, factor=[1, 1, 1], input_dim=5, scm="linear", flow_num=1, beta=0.01, lr=0.01)
        cfg["lambda"] = 10.0
        spec = orc.tabular_spec(cfg, [1, 1, 3], [2, 3, 0, 1, 4])
        A = orc.i_b_inv(orc.tabular_B("adult"))
        params = orc.init_params(spec, 1)
        x, y, noise = orc.synth_tabular("adult", 64)
        lo, hi = D.shard_rows(64)
        assert (hi - lo) * world == 64 and lo == rank * 32
        # local mean-of-shard gradient, flattened into an "arena" with a gap that must not be touched
        leaves = {k: v.clone().requires_grad_(True) for k, v in params.items()}
        loss, logs, _ = orc.step_losses(leaves, spec, A, x[lo:hi], y[lo:hi], noise[lo:hi])
        gl = torch.autograd.grad(loss, list(leaves.values()))
        flat = torch.cat([g.reshape(-1) for g in gl])
        arena = torch.cat([flat, torch.full((7,), 123.0)])
        scale = D.allreduce_arena(arena, [(0, 40), (40, flat.numel() - 40)], bucket_floats=16)
        assert scale == 1.0 / world
        assert torch.equal(arena[-7:], torch.full((7,), 123.0))
        # full-batch gradient from one process
        leaves2 = {k: v.clone().requires_grad_(True) for k, v in params.items()}
        loss2, logs2, _ = orc.step_losses(leaves2, spec, A, x, y, noise)
        gl2 = torch.autograd.grad(loss2, list(leaves2.values()))
        flat2 = torch.cat([g.reshape(-1) for g in gl2])
        err = float((arena[:-7] * scale - flat2).norm() / flat2.norm())
        row = torch.tensor([float(logs[k]) for k in ("loss", "recon", "KL", "alignment")])
        D.allreduce_mean_(row)
        row2 = torch.tensor([float(logs2[k]) for k in ("loss", "recon", "KL", "alignment")])
        lerr = float((row - row2).abs().max() / row2.abs().max())
        q.put((rank, err, lerr))
    finally:
        dist.destroy_process_group()


def test_gradient_exchange_world2_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in procs]
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    for rank, err, lerr in res:
        assert err < 1e-5, (rank, err)          # mean of shard means == full-batch mean (equal shards)
        assert lerr < 1e-5, (rank, lerr)


def test_shard_rows_rejects_ragged_split():
    from cdgvae_b200 import dist as D
    assert D.world() == 1 and D.shard_rows(10) == (0, 10)
    assert D.allreduce_arena(torch.ones(4), [(0, 4)]) == 1.0
