"""Diagnostic: gradient error of every parameter against the oracle at batch 2048 (the pre-split / planes routes), per switch."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from oracle import cdgvae_oracle as orc
from cdgvae_b200.modules.model import CDGVAE
from cdgvae_b200.modules import train as T

def rel(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return float((a - b).norm() / (b.norm() + 1e-30))

B = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
for scm in ("linear", "nonlinear"):
    cfg = dict(node=4, scm=scm, flow_num=1, inverse_loop=100, factor=[1, 1, 2], image_size=64, batch_size=B, lr=1e-3, beta=0.1, seed=1)
    cfg["lambda"] = 5.0
    Bm, mask = orc.pendulum_B(4), orc.pendulum_masks(64)
    spec = orc.pendulum_spec(cfg, mask)
    torch.manual_seed(1)
    model = CDGVAE(Bm, mask, cfg, "cpu").to("cuda")
    opt = torch.optim.Adam(model.parameters(), lr=cfg["lr"])
    x, y, noise = orc.synth_pendulum(B, 64, 4, 1234, 4321)
    model.noise_fn = lambda n, d: noise
    logs, xhat = T.train_CDGVAE([(x, y)], model, cfg, opt, "cuda")
    oparams = orc.init_params(spec, 1)
    ol, og, oo = orc.train_step(oparams, orc.new_adam_state(oparams), spec, orc.i_b_inv(Bm), x, y, noise)
    # fp64 oracle for the noise floor of the fp32 oracle itself
    op64 = {k: v.double() for k, v in oparams.items()}
    print(scm, "env", {k: v for k, v in os.environ.items() if k.startswith("CDG_P")})
    for n, p in model.named_parameters():
        print(f"   {n:22s} {rel(p.grad, og[n]):.2e}")
