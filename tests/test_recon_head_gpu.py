"""The fused reconstruction head, ELEMENT BY ELEMENT: d loss / d pre as the training step leaves it in its workspace (fp32, or
bf16 hi + lo planes on the pre-split route) against a float64 evaluation of model.py:287 / train.py:175 from the step's own
a2, the weights and the target.  The per-tensor 1e-4 gradient checks elsewhere average over millions of elements: a race in
the TMA-staged epilogue once left ~30 wrong elements in 8 million (stale 32-byte tails of target rows) and every one of those
checks stayed green.  Here no single element may be off by more than 5e-5 of the largest gradient entry (the operands of the
product carry 16 mantissa bits each; a stale target is off by ~1), over several steps."""
import pytest
import torch

from oracle import cdgvae_oracle as orc

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("batch,scm", [(2048, "linear"), (8192, "nonlinear"), (4096 + 256, "linear")])
def test_recon_gradient_elementwise(batch, scm):
    from cdgvae_b200 import _lib
    from cdgvae_b200.modules.model import CDGVAE
    from cdgvae_b200.modules import train as T
    B, P, H = batch, 12288, 300
    cfg = dict(node=4, scm=scm, flow_num=1, inverse_loop=100, factor=[1, 1, 2], image_size=64, batch_size=B, lr=1e-3, beta=0.1, seed=1)
    cfg["lambda"] = 5.0
    Bm, mask = orc.pendulum_B(4), orc.pendulum_masks(64)
    torch.manual_seed(1)
    model = CDGVAE(Bm, mask, cfg, "cpu").to("cuda")
    opt = torch.optim.Adam(model.parameters(), lr=cfg["lr"])
    bands = [(0, 3840), (3840, 9792), (9792, 12288)]
    elu = lambda t: torch.where(t > 0, t, torch.exp(t) - 1)
    for step in range(3):
        x, y, noise = orc.synth_pendulum(B, 64, 4, 1234 + step, 4321 + step)
        model.noise_fn = lambda n, d: noise
        sd = {k: v.detach().double().clone() for k, v in model.state_dict().items()}      # the parameters this step starts from
        T.train_CDGVAE([(x, y)], model, cfg, opt, "cuda")
        torch.cuda.synchronize()
        plan = model._get_plan()
        ws = model._workspace.view(torch.float32)
        off = lambda w: _lib.lib().cdg_pendulum_workspace_offset(plan, B, 0, w)
        raw = ws[off(0): off(0) + B * P]
        if off(100) == 1:                                   # bf16 planes: hi, then lo
            u16 = raw.view(torch.bfloat16)
            g = (u16[: B * P].double() + u16[B * P: 2 * B * P].double()).view(B, P)
        else:
            g = raw.double().view(B, P)
        tol = 5e-5
        xd = x.cuda().double().view(B, P)
        for k, (lo, hi) in enumerate(bands):
            a2 = ws[off(5 + k): off(5 + k) + B * H].double().view(B, H)
            pre = a2 @ sd[f"decoder.{k}.4.weight"][lo:hi].t() + sd[f"decoder.{k}.4.bias"][lo:hi]
            t = torch.tanh(pre)
            ref = (t - xd[:, lo:hi]) * (1 - t * t) / B
            err = (g[:, lo:hi] - ref).abs()
            worst = float(err.max() / ref.abs().max())
            assert worst < tol, (batch, scm, step, k, worst, int((err > tol * ref.abs().max()).sum()))
