"""Diagnostic: d loss / d pre and a2 as the step left them in its workspace, per decoder, against a float64 evaluation of the
same quantities from the model's own parameters (batch 2048, linear SCM)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from oracle import cdgvae_oracle as orc
from cdgvae_b200 import _lib
from cdgvae_b200.modules.model import CDGVAE
from cdgvae_b200.modules import train as T

B = 2048
scm = sys.argv[1] if len(sys.argv) > 1 else "linear"
cfg = dict(node=4, scm=scm, flow_num=1, inverse_loop=100, factor=[1, 1, 2], image_size=64, batch_size=B, lr=1e-3, beta=0.1, seed=1)
cfg["lambda"] = 5.0
Bm, mask = orc.pendulum_B(4), orc.pendulum_masks(64)
torch.manual_seed(1)
model = CDGVAE(Bm, mask, cfg, "cpu").to("cuda")
sd = {k: v.detach().double().clone() for k, v in model.state_dict().items()}
opt = torch.optim.Adam(model.parameters(), lr=0.0)            # lr 0: parameters stay put
x, y, noise = orc.synth_pendulum(B, 64, 4, 1234, 4321)
model.noise_fn = lambda n, d: noise
logs, xhat = T.train_CDGVAE([(x, y)], model, cfg, opt, "cuda")
torch.cuda.synchronize()
plan = model._get_plan()
ws = model._workspace.view(torch.float32)
off = lambda w: _lib.lib().cdg_pendulum_workspace_offset(plan, B, 0, w)
P, H = 12288, 300
planes = os.environ.get("CDG_PK", "1") != "0" and os.environ.get("CDG_PS", "1") != "0"
pre_raw = ws[off(0): off(0) + B * P]
if planes:
    u16 = pre_raw.view(torch.bfloat16)
    g = (u16[: B * P].double() + u16[B * P: 2 * B * P].double()).view(B, P)
else:
    g = pre_raw.double().view(B, P)
xd = x.cuda().double().view(B, P)
bands = [(0, 3840), (3840, 9792), (9792, 12288)]
elu = lambda t: torch.where(t > 0, t, torch.exp(t) - 1)
for k, (lo, hi) in enumerate(bands):
    a2 = ws[off(5 + k): off(5 + k) + B * H].double().view(B, H)
    a1 = ws[off(1 + k): off(1 + k) + B * H].double().view(B, H)
    a2_ref = elu(a1 @ sd[f"decoder.{k}.2.weight"].t() + sd[f"decoder.{k}.2.bias"])
    pre = a2 @ sd[f"decoder.{k}.4.weight"][lo:hi].t() + sd[f"decoder.{k}.4.bias"][lo:hi]
    t = torch.tanh(pre)
    ref = (t - xd[:, lo:hi]) * (1 - t * t) / B
    d = g[:, lo:hi]
    err = (d - ref)
    print(f"dec {k}: a2 rel {float((a2 - a2_ref).norm() / a2_ref.norm()):.2e}   d_pre rel {float(err.norm() / ref.norm()):.2e}  "
          f"mean signed err / mean |ref| {float(err.mean() / ref.abs().mean()):+.2e}  max |err| / max |ref| {float(err.abs().max() / ref.abs().max()):.2e}  "
          f"colsum rel {float((d.sum(0) - ref.sum(0)).norm() / ref.sum(0).norm()):.2e}")
    bad = (err.abs() > 1e-3 * ref.abs().max()).nonzero()
    print(f"    {bad.shape[0]} elements off by more than 1e-3 of max |ref|; first 24 (row, col-in-band, got*B, ref*B, x, t):")
    for r, cidx in bad[:24].tolist():
        print(f"      row {r:5d} (pair {r // 256}, cta {(r // 128) % 2}, q {(r % 128) // 32}, lane {r % 32})  col {cidx:5d} (tile {cidx // 256}, "
              f"half {(cidx % 256) // 128}, chunk {(cidx % 128) // 32}, in-chunk {cidx % 32})  got {float(d[r, cidx]) * B:+.4f} ref {float(ref[r, cidx]) * B:+.4f} "
              f"x {float(xd[r, lo + cidx]):+.4f} t {float(t[r, cidx]):+.4f}")
