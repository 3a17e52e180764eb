"""CelebA CDG-VAE step throughput (BASELINE configs[4]): samples/s of train_CDGVAE's loop body at a given per-GPU batch,
inputs resident on the device, CUDA-event timing.  Under torchrun (one rank per GPU) the batch is per rank, gradients are
all-reduced over NCCL (weak scaling; BatchNorm statistics stay per shard) and the time is the max over ranks.

    python tools/celeba_bench.py [batch] [steps] [gemm_mode]
    python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 tools/celeba_bench.py 16 10"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from cdgvae_b200 import _lib
from cdgvae_b200.celeba.module.model import CDGVAE
from cdgvae_b200.celeba.module.train import train_CDGVAE

def main():
    batch = int(sys.argv[1]) if len(sys.argv) > 1 else 16
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
    mode = sys.argv[3] if len(sys.argv) > 3 else "auto"
    import torch.distributed as dist
    world, rank, local = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("NCCL_DEBUG", "WARN")
        dist.init_process_group("nccl", device_id=dev)
    cfg = dict(node=6, latent_dim=6, scm="linear", flow_num=1, inverse_loop=100, beta=0.1, lr=1e-3, batch_size=batch,
               pretrained=False, gemm_mode=mode)
    cfg["lambda"] = 5.0
    g = torch.Generator(device=dev).manual_seed(1234 + rank)
    x = torch.cat([torch.rand(batch, 128, 128, 3, device=dev, generator=g),
                   (torch.rand(batch, 128, 128, 5, device=dev, generator=g) < 0.5).float()], -1)
    y = (torch.rand(batch, 6, device=dev, generator=g) < 0.5).float()
    B = torch.zeros(6, 6); B[0, 2] = B[0, 3] = B[0, 5] = 1; B[0, 4] = B[1, 4] = 0.5
    if os.environ.get("CDG_CELEBA_STREAMS"):
        _lib.lib().cdg_celeba_generator_streams(int(os.environ["CDG_CELEBA_STREAMS"]))
    torch.manual_seed(1)
    model = CDGVAE(B, torch.split(x[..., 3:], 1, dim=-1), cfg, dev)
    opt = torch.optim.Adam(model.parameters(), lr=1e-3)
    noise = torch.randn(batch, 6, device=dev)
    model.noise_fn = lambda b, d: noise
    train_CDGVAE([(x, y)] * 6, model, cfg, opt, dev)
    torch.cuda.synchronize()
    n0 = _lib.lib().cdg_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    logs, _ = train_CDGVAE([(x, y)] * steps, model, cfg, opt, dev)
    e1.record()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / steps], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t)
    if rank == 0:
        print(json.dumps({"n_gpus": world, "samples_per_s_total": world * batch / ms * 1e3, "workload": "celeba CDG-VAE train step", "batch": batch, "gemm_mode": mode, "ms_per_step": ms,
                      "samples_per_s": batch / ms * 1e3, "launches_per_step": (_lib.lib().cdg_launch_count() - n0) / steps,
                      "algorithmic_tflops": 47.6e9 * batch / ms / 1e9, "loss": logs["loss"][-1],
                      "workspace_gb": model._workspace.numel() / 1e9}))
    if world > 1:
        dist.destroy_process_group()

if __name__ == "__main__":
    main()
