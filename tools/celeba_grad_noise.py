"""Diagnostic: gradient error of the CelebA step vs the fp64 oracle, for the tcgen05 3xTF32 and the SIMT fp32 GEMM modes,
next to the fp32 oracle's own error (the reference arithmetic's noise floor)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import cdgvae_oracle as orc, celeba_oracle as corc
from cdgvae_b200.celeba.module.model import CDGVAE
from cdgvae_b200.celeba.module.train import train_CDGVAE

def rel(a, b):
    a, b = a.detach().double().cpu().reshape(-1), b.detach().double().cpu().reshape(-1)
    return float((a - b).norm() / (b.norm() + 1e-30))

batch = int(sys.argv[1]) if len(sys.argv) > 1 else 2
for seed in (1234, 1235, 1236):
    res = {}
    for mode in ("auto", "simt"):
        cfg = dict(node=6, latent_dim=6, scm="linear", flow_num=1, inverse_loop=100, beta=0.1, lr=1e-3, seed=1, batch_size=batch,
                   pretrained=False, gemm_mode=mode)
        cfg["lambda"] = 5.0
        x, y, n1, n2 = corc.synth_celeba(batch, seed, seed + 1)
        masks = torch.split(x[..., 3:], 1, dim=-1)
        torch.manual_seed(1)
        model = CDGVAE(corc.celeba_B(), masks, cfg, "cpu").to("cuda")
        opt = torch.optim.Adam(model.parameters(), lr=1e-3)
        q = [n1, n2]
        model.noise_fn = lambda b, d: q.pop(0)
        logs, xhat = train_CDGVAE([(x, y)], model, cfg, opt, "cuda")
        res[mode] = {n: p.grad.clone() for n, p in model.named_parameters() if p.requires_grad}
    spec = corc.CelebaSpec(cfg)
    A = orc.i_b_inv(corc.celeba_B())
    out = {}
    for dt in (torch.float32, torch.float64):
        st = {k: (v.to(dt) if v.dtype.is_floating_point else v) for k, v in corc.init_state(cfg, 1).items()}
        _, g, _ = corc.train_step(st, corc.new_adam_state(st), spec, A.to(dt), x.to(dt), y.to(dt), [m.to(dt) for m in masks], n1.to(dt), n2.to(dt))
        out[dt] = g
    for n in out[torch.float64]:
        print(f"seed {seed} {n:18s} oracle32 {rel(out[torch.float32][n], out[torch.float64][n]):.2e}  tc3x {rel(res['auto'][n], out[torch.float64][n]):.2e}"
              f"  simt {rel(res['simt'][n], out[torch.float64][n]):.2e}")
