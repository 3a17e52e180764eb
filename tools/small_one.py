"""One short run of the pendulum step at the reference's batch 128 (for `ncu --metrics gpu__time_duration.sum`)."""
import sys
sys.path.insert(0, ".")
import torch
import synthetic_inputs as syn
from cdgvae_b200.modules.model import CDGVAE
from cdgvae_b200.modules.train import train_CDGVAE

B = int(sys.argv[1]) if len(sys.argv) > 1 else 128
cfg = dict(node=4, scm="linear", flow_num=1, inverse_loop=100, factor=[1, 1, 2], image_size=64, batch_size=B, lr=1e-3, beta=0.1)
cfg["lambda"] = 5.0
torch.manual_seed(1)
model = CDGVAE(syn.pendulum_B(4), syn.pendulum_masks(64), cfg, "cpu").to("cuda")
opt = torch.optim.Adam(model.parameters(), lr=1e-3)
x, y, noise = syn.synth_pendulum(B)
data = [(x.cuda(), y.cuda())] * 6
train_CDGVAE(data, model, cfg, opt, "cuda")
torch.cuda.synchronize()
