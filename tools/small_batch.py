"""Step time at the reference's own batch sizes (pendulum bs 128, main.py:96; tabular bs 256, tabular/main.py:88)."""
import sys
import time

import torch

sys.path.insert(0, ".")
from cdgvae_b200.modules.model import CDGVAE
from cdgvae_b200.modules.train import train_CDGVAE
from oracle import cdgvae_oracle as orc

cfg = dict(node=4, scm="linear", flow_num=1, inverse_loop=100, factor=[1, 1, 2], image_size=64, batch_size=128, lr=1e-3, beta=0.1)
cfg["lambda"] = 5.0
torch.manual_seed(1)
model = CDGVAE(orc.pendulum_B(4), orc.pendulum_masks(64), cfg, "cpu").to("cuda")
opt = torch.optim.Adam(model.parameters(), lr=1e-3)
for B in (128, 1024):
    x, y, noise = orc.synth_pendulum(B)
    data = [(x.cuda(), y.cuda())] * 59          # one reference epoch: 7,500 images / 128
    train_CDGVAE(data[:5], model, cfg, opt, "cuda")
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    train_CDGVAE(data, model, cfg, opt, "cuda")
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print(f"pendulum train_CDGVAE batch {B}: {1e3 * dt / 59:.3f} ms/step wall, {B * 59 / dt:.0f} samples/s (device-resident batches, CPU noise draw)")
    hx = [(x.pin_memory(), y.pin_memory())] * 59
    t0 = time.perf_counter()
    train_CDGVAE(hx, model, cfg, opt, "cuda")
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print(f"pendulum train_CDGVAE batch {B}: {1e3 * dt / 59:.3f} ms/step wall, {B * 59 / dt:.0f} samples/s (host batches)")

from collections import namedtuple
from cdgvae_b200.tabular.modules import model as TM, train as TT
cfg = dict(dataset="adult", scm="linear", flow_num=1, inverse_loop=100, lr=0.01, beta=0.01, node=3, factor=[1, 1, 1], input_dim=5)
cfg["lambda"] = 10.0
torch.manual_seed(1)
tm = TM.CDGVAE(orc.tabular_B("adult"), [1, 1, 3], cfg, "cpu").to("cuda")
topt = torch.optim.Adam(tm.parameters(), lr=0.01)
DS = namedtuple("DS", ["flatten_topology"])
x, y, _ = orc.synth_tabular("adult", 256)
data = [(x.cuda(), y.cuda())] * 157                # one reference epoch: 40,000 rows / 256
TT.train_CDGVAE(DS([2, 3, 0, 1, 4]), data[:5], tm, cfg, topt, "cuda")
torch.cuda.synchronize()
t0 = time.perf_counter()
TT.train_CDGVAE(DS([2, 3, 0, 1, 4]), data, tm, cfg, topt, "cuda")
torch.cuda.synchronize()
dt = time.perf_counter() - t0
print(f"tabular adult train_CDGVAE batch 256: {1e3 * dt / 157:.3f} ms/step wall, {256 * 157 / dt:.0f} rows/s")
