import sys, torch
sys.path.insert(0, ".")
from cdgvae_b200.modules.model import CDGVAE
from cdgvae_b200.modules.train import train_CDGVAE
from oracle import cdgvae_oracle as orc
cfg = dict(node=4, scm="linear", flow_num=1, inverse_loop=100, factor=[1, 1, 2], image_size=64, batch_size=128, lr=1e-3, beta=0.1)
cfg["lambda"] = 5.0
torch.manual_seed(1)
model = CDGVAE(orc.pendulum_B(4), orc.pendulum_masks(64), cfg, "cpu").to("cuda")
model.use_graphs = False
opt = torch.optim.Adam(model.parameters(), lr=1e-3)
x, y, noise = orc.synth_pendulum(128)
train_CDGVAE([(x.cuda(), y.cuda())] * 4, model, cfg, opt, "cuda")
torch.cuda.synchronize()
print("ok")
