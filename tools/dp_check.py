"""Data-parallel equivalence on real GPUs (run under torchrun, one rank per GPU): every rank first trains the model on the
FULL global batches in a single-process setting, then the ranks train on their contiguous row shards with the gradient
all-reduce (CUDA-graph replay with the NCCL call captured for the tabular step; events + communication stream for the
pendulum step); the (all-reduced, averaged) gradients of the last step must agree at 1e-4 per tensor and the updated parameters are reported
(summation order differs; tiny-gradient elements move by +-lr under Adam for any implementation, so the
pendulum case runs ONE step and judges gradients; the tabular case, whose trajectory is well conditioned, runs eight).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 tools/dp_check.py
"""
import os
import sys
from collections import namedtuple

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402
import synthetic_inputs as syn  # noqa: E402


def rel(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return float((a - b).norm() / (b.norm() + 1e-30))


def tabular(dev, world, rank, steps=8, rows=4096):
    from cdgvae_b200.tabular.modules import model as M, train as T
    cfg = dict(dataset="adult", scm="linear", flow_num=1, inverse_loop=100, lr=0.01, beta=0.01, node=3, factor=[1, 1, 1], input_dim=5)
    cfg["lambda"] = 10.0
    DS = namedtuple("DS", ["flatten_topology"])
    data = [syn.synth_tabular("adult", rows, 100 + s, 200 + s) for s in range(steps)]

    def run(shard):
        torch.manual_seed(1)
        m = M.CDGVAE(syn.tabular_B("adult"), [1, 1, 3], cfg, "cpu").to(dev)
        opt = torch.optim.Adam(m.parameters(), lr=cfg["lr"])
        lo, hi = shard
        q = [n[lo:hi] for _, _, n in data]
        m.noise_fn = lambda b, d: q.pop(0)
        T.train_CDGVAE(DS([2, 3, 0, 1, 4]), [(x[lo:hi], y[lo:hi]) for x, y, _ in data], m, cfg, opt, dev)
        return ({k: v.detach().clone() for k, v in m.state_dict().items()},
                {k: p.grad.detach().clone() for k, p in m.named_parameters() if p.grad is not None}), m
    return run, rows


def pendulum(dev, world, rank, steps=1, rows=4096):
    from cdgvae_b200.modules.model import CDGVAE
    from cdgvae_b200.modules import train as T
    cfg = dict(node=4, scm="nonlinear", flow_num=1, inverse_loop=100, factor=[1, 1, 2], image_size=64, batch_size=rows, lr=1e-3, beta=0.1, seed=1)
    cfg["lambda"] = 5.0
    data = [syn.synth_pendulum(rows, 64, 4, 300 + s, 400 + s) for s in range(steps)]

    def run(shard):
        torch.manual_seed(1)
        m = CDGVAE(syn.pendulum_B(4), syn.pendulum_masks(64), cfg, "cpu").to(dev)
        opt = torch.optim.Adam(m.parameters(), lr=cfg["lr"])
        lo, hi = shard
        q = [n[lo:hi] for _, _, n in data]
        m.noise_fn = lambda b, d: q.pop(0)
        T.train_CDGVAE([(x[lo:hi], y[lo:hi]) for x, y, _ in data], m, cfg, opt, dev)
        return ({k: v.detach().clone() for k, v in m.state_dict().items()},
                {k: p.grad.detach().clone() for k, p in m.named_parameters() if p.grad is not None}), m
    return run, rows


def main():
    world, rank, local = int(os.environ["WORLD_SIZE"]), int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = f"cuda:{local}"
    cases = {"tabular_adult": tabular(dev, world, rank), "pendulum_nonlinear": pendulum(dev, world, rank)}
    full = {name: run((0, rows))[0] for name, (run, rows) in cases.items()}          # before the process group exists: world = 1
    dist.init_process_group("nccl", device_id=torch.device(dev))
    ok = True
    for name, (run, rows) in cases.items():
        per = rows // world
        (sd, gr), m = run((rank * per, (rank + 1) * per))
        fsd, fgr = full[name]
        worst = max((rel(sd[k], fsd[k]), k) for k in sd if fsd[k].numel() > 0)
        # the gradient arena holds the SUM over ranks after the exchange (1 / world is folded into the Adam kernel)
        gworst = max((rel(gr[k] / world, fgr[k]), k) for k in gr if fgr[k].numel() > 2)
        graphs = len(getattr(m, "_graphs", {}) or {})
        print(f"[rank {rank}] {name}: gradients sharded-vs-full worst {gworst[0]:.2e} ({gworst[1]}), updated parameters worst "
              f"{worst[0]:.2e} ({worst[1]}), graphs {graphs}, dp_graphs {getattr(m, 'dp_graphs', True)}", flush=True)
        ok = ok and gworst[0] < 1e-4
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
