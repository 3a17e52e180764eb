"""CPU oracle for the CelebA CDG-VAE training step.  TEST INFRASTRUCTURE ONLY.

A from-scratch restatement (torch CPU tensors, functional ops + autograd) of

    celeba/module/model.py:106-218     CDGVAE (frozen ResNet-18 encoder, two posteriors, 5 SAGAN generators)
    celeba/module/sagan.py:74-210      NoiseInjection / GenIniBlock / GenBlock / Generator
    celeba/module/train.py:10-76       train_CDGVAE
    torchvision.models.resnet18        (third-party; torchvision 0.26 in this image; restated below)
    torch.nn.utils.spectral_norm       (third-party; one power iteration per training-mode forward)

Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs may import this file; the product
package never does.

Parity pinning: the reference ships no tests or golden vectors; tests/golden/make_golden.py imports the
UNMODIFIED reference from /root/reference (with torchvision's `pretrained=True` download replaced by random
initialisation, because there is no network and no cached checkpoint), runs celeba/module/train.py's
train_CDGVAE on synthetic inputs, cross-checks this file step by step and writes tests/golden/celeba_*.json.

Reference behaviours restated literally:
  * `self.decoder` is a plain Python list, so the generators are never optimised and stay in train mode;
    only `encoder.fc.{weight,bias}` and `flows.i.p` are trainable;
  * NoiseInjection.weight and Self_Attn.sigma are initialised to 0 and never trained, so both layers return
    their input exactly (`x + 0 * finite`); they are asserted to be zero and skipped;
  * `Generator.apply(init_weights)` orthogonalises the *derived* `.weight` attribute of the spectral-norm
    wrapped layers, which the next training-mode forward overwrites with weight_orig / sigma: the effective
    initial weights are the default nn.Linear / nn.Conv2d draws, the biases are 0;
  * the frozen ResNet-18 runs in train mode: batch statistics in the forward, running statistics updated on
    each of the two encode() calls of a forward;
  * both noise draws are shaped [B, node]; KL2 subtracts node.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Tuple

import torch
import torch.nn.functional as F

from . import cdgvae_oracle as base

Tensor = torch.Tensor

# decoder k reads these latents (celeba/module/model.py:190-195); the 5th reads epsilon2
DEC_INPUTS = [[0, 2], [0, 3], [0, 4], [0, 1, 5]]
GEN_BLOCKS = [(512, 512), (512, 256), (256, 128), (128, 64), (64, 32)]   # sagan.py:163-178 with conv_dim = 32, 128 px
RESNET_LAYERS = [(64, 1), (128, 2), (256, 2), (512, 2)]                   # torchvision resnet18: (planes, stride)


class CelebaSpec:
    def __init__(self, config: dict):
        self.node = config["node"]
        self.latent_dim = config["latent_dim"]
        assert self.node == 6, "the decoder wiring of celeba/module/model.py:190-195 needs 6 causal latents"
        self.scm = config["scm"]
        self.flow_num = config.get("flow_num", 1)
        self.beta, self.lam, self.lr = config["beta"], config["lambda"], config.get("lr", 1e-3)
        self.betas, self.eps, self.weight_decay = (0.9, 0.999), 1e-8, 0.0
        self.z_dims = [len(i) for i in DEC_INPUTS] + [self.latent_dim]


import os as _os
import sys as _sys

_sys.path.insert(0, _os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))))
from synthetic_inputs import celeba_B, synth_celeba  # noqa: E402,F401  (shared with bench.py; see synthetic_inputs.py)


# --------------------------------------------------------------------------------------
# Same-seed initialisation: creation order of celeba/module/model.py:106-151
# --------------------------------------------------------------------------------------
def init_state(config: dict, seed: int = 1) -> Dict[str, Tensor]:
    """State created by `torch.manual_seed(seed); CDGVAE(B, mask, config, 'cpu')` with a randomly initialised
    ResNet-18: registered state (`encoder.*`, `flows.*`) plus the unregistered generators under `decoder.k.*`.
    torchvision / nn.Linear / nn.Conv2d / spectral_norm / orthogonal_ supply the init laws and RNG consumption."""
    import torch.nn as nn
    import torchvision
    from torch.nn.utils import spectral_norm
    torch.manual_seed(seed)
    sd: Dict[str, Tensor] = {}
    enc = torchvision.models.resnet18(weights=None)                             # model.py:117
    enc.fc = nn.Linear(512, config["node"] * 2 + config["latent_dim"] * 2)      # model.py:118-120
    for k, v in enc.state_dict().items():
        sd["encoder." + k] = v.detach().clone()
    for i in range(config["node"]):
        if config["scm"] == "linear":
            sd[f"flows.{i}.p"] = torch.rand([2]) * 0.1
        else:
            for nm in ("w", "b", "u"):
                for j in range(config["flow_num"]):
                    sd[f"flows.{i}.{nm}.{j}"] = torch.randn(1, 1) * 0.1
    spec = CelebaSpec(config)
    for k, zd in enumerate(spec.z_dims):
        made = []        # (prefix, module) in construction order; init_weights visits them in the same order

        def sn(prefix, mod):
            mod = spectral_norm(mod)
            made.append((prefix, mod))
            return mod

        sn("block0.snlinear0", nn.Linear(zd, 512 * 16))                          # sagan.py:91
        noise = {"block0.noise0.weight": torch.zeros(1, 512, 4, 4)}              # sagan.py:93 (size 4)
        bns = {}
        for b, (ci, co) in enumerate(GEN_BLOCKS, 1):                             # sagan.py:107-121
            sn(f"block{b}.conv_1", nn.Conv2d(ci, co, 3, 1, 1))
            sn(f"block{b}.conv_2", nn.Conv2d(co, co, 3, 1, 1))
            size = {1: 8, 2: 16}.get(b, 1)                                       # sagan.py:163-164, :174-178
            noise[f"block{b}.noise1.weight"] = torch.zeros(1, co, size, size)
            noise[f"block{b}.noise2.weight"] = torch.zeros(1, co, size, size)
            sn(f"block{b}.conv_0", nn.Conv2d(ci, co, 1, 1, 0))
            bns[f"block{b}.bn1"], bns[f"block{b}.bn2"] = ci, co
            if b == 3:                                                           # sagan.py:175-176
                for nm, (a, o) in (("theta", (128, 16)), ("phi", (128, 16)), ("g", (128, 64)), ("attn", (64, 128))):
                    sn(f"self_attn1.snconv1x1_{nm}", nn.Conv2d(a, o, 1, 1, 0))
        bns["bn"] = 32
        sn("toRGB", nn.Conv2d(32, 3, 3, 1, 1))
        # `self.apply(init_weights)` (sagan.py:190): children first, registration order.  orthogonal_ lands on the
        # derived `.weight` tensor (not on weight_orig) but consumes the RNG; the bias is filled with 0.
        order = ["block0.snlinear0"]
        for b in range(1, 6):
            order += [f"block{b}.conv_1", f"block{b}.conv_2", f"block{b}.conv_0"]
            if b == 3:
                order += [f"self_attn1.snconv1x1_{nm}" for nm in ("theta", "phi", "g", "attn")]
        order.append("toRGB")
        mods = dict(made)
        for name in order:
            m = mods[name]
            nn.init.orthogonal_(m.weight)
            m.bias.data.fill_(0.0)
        pre = f"decoder.{k}."
        for name, m in made:
            sd[pre + name + ".bias"] = m.bias.detach().clone()
            sd[pre + name + ".weight_orig"] = m.weight_orig.detach().clone()
            sd[pre + name + ".weight_u"] = m.weight_u.detach().clone()
            sd[pre + name + ".weight_v"] = m.weight_v.detach().clone()
        for n, t in noise.items():
            sd[pre + n] = t
        sd[pre + "self_attn1.sigma"] = torch.zeros(1)
        for n, c in bns.items():
            sd[pre + n + ".weight"], sd[pre + n + ".bias"] = torch.ones(c), torch.zeros(c)
            sd[pre + n + ".running_mean"], sd[pre + n + ".running_var"] = torch.zeros(c), torch.ones(c)
            sd[pre + n + ".num_batches_tracked"] = torch.tensor(0)
    return sd


TRAINABLE = ("encoder.fc.weight", "encoder.fc.bias")


def trainable_names(state: Dict[str, Tensor]) -> List[str]:
    """celeba/module/model.py:121-125 + the registered flows: everything else is frozen or unregistered."""
    return [k for k in state if k in TRAINABLE or k.startswith("flows.")]


# --------------------------------------------------------------------------------------
# Layers
# --------------------------------------------------------------------------------------
def batch_norm_train(state, prefix: str, x: Tensor, momentum: float = 0.1, eps: float = 1e-5) -> Tensor:
    """nn.BatchNorm2d in training mode: batch statistics normalise, running statistics (unbiased variance) and
    num_batches_tracked are updated in place."""
    n = x.numel() // x.shape[1]
    mean = x.mean(dim=(0, 2, 3))
    var = x.var(dim=(0, 2, 3), unbiased=False)
    with torch.no_grad():
        state[prefix + ".running_mean"].mul_(1 - momentum).add_(momentum * mean.detach())
        state[prefix + ".running_var"].mul_(1 - momentum).add_(momentum * var.detach() * n / max(n - 1, 1))
        state[prefix + ".num_batches_tracked"].add_(1)
    w, b = state[prefix + ".weight"], state[prefix + ".bias"]
    xh = (x - mean[None, :, None, None]) / torch.sqrt(var[None, :, None, None] + eps)
    return xh * w[None, :, None, None] + b[None, :, None, None]


def resnet18_features(state, x: Tensor, pre: str = "encoder.") -> Tensor:
    """torchvision ResNet._forward_impl up to (and including) avgpool + flatten, BasicBlock.forward."""
    h = F.conv2d(x, state[pre + "conv1.weight"], None, 2, 3)
    h = F.relu(batch_norm_train(state, pre + "bn1", h))
    h = F.max_pool2d(h, 3, 2, 1)
    for li, (planes, stride) in enumerate(RESNET_LAYERS, 1):
        for bi in range(2):
            p = f"{pre}layer{li}.{bi}."
            s = stride if bi == 0 else 1
            o = F.conv2d(h, state[p + "conv1.weight"], None, s, 1)
            o = F.relu(batch_norm_train(state, p + "bn1", o))
            o = F.conv2d(o, state[p + "conv2.weight"], None, 1, 1)
            o = batch_norm_train(state, p + "bn2", o)
            if p + "downsample.0.weight" in state:
                idn = F.conv2d(h, state[p + "downsample.0.weight"], None, s, 0)
                idn = batch_norm_train(state, p + "downsample.1", idn)
            else:
                idn = h
            h = F.relu(o + idn)
    return h.mean(dim=(2, 3))


def spectral_weight(state, prefix: str) -> Tensor:
    """torch.nn.utils.spectral_norm in training mode: one power iteration updates u and v in place, then
    weight = weight_orig / (u . W v)."""
    w = state[prefix + ".weight_orig"]
    wm = w.reshape(w.shape[0], -1)
    u, v = state[prefix + ".weight_u"], state[prefix + ".weight_v"]
    with torch.no_grad():
        v.copy_(F.normalize(torch.mv(wm.t(), u), dim=0, eps=1e-12))
        u.copy_(F.normalize(torch.mv(wm, v), dim=0, eps=1e-12))
    sigma = torch.dot(u, torch.mv(wm, v))
    return w / sigma


def gen_block(state, p: str, x: Tensor) -> Tensor:
    """GenBlock.forward, sagan.py:123-140 (noise layers are exact no-ops, see the header)."""
    for nz in ("noise1", "noise2"):
        assert not bool(state[p + nz + ".weight"].any())
    h = F.relu(batch_norm_train(state, p + "bn1", x))
    h = F.interpolate(h, scale_factor=2, mode="nearest")
    h = F.conv2d(h, spectral_weight(state, p + "conv_1"), state[p + "conv_1.bias"], 1, 1)
    h = F.relu(batch_norm_train(state, p + "bn2", h))
    h = F.conv2d(h, spectral_weight(state, p + "conv_2"), state[p + "conv_2.bias"], 1, 1)
    x0 = F.interpolate(x, scale_factor=2, mode="nearest")
    x0 = F.conv2d(x0, spectral_weight(state, p + "conv_0"), state[p + "conv_0.bias"], 1, 0)
    return h + x0


def generator(state, k: int, z: Tensor) -> Tensor:
    """Generator.forward, sagan.py:192-210, image_size 128 (NCHW output in [-1, 1])."""
    p = f"decoder.{k}."
    assert not bool(state[p + "block0.noise0.weight"].any()) and not bool(state[p + "self_attn1.sigma"].any())
    h = F.linear(z, spectral_weight(state, p + "block0.snlinear0"), state[p + "block0.snlinear0.bias"])
    h = h.view(-1, 512, 4, 4)                                                    # sagan.py:97
    for b in range(1, 6):
        h = gen_block(state, f"{p}block{b}.", h)
        if b == 3:
            # Self_Attn (sagan.py:46-73): out = x + sigma * attn_g with sigma == 0, i.e. x exactly; the only lasting
            # effect of the layer is the power iteration of its four spectral-norm convolutions
            for nm in ("theta", "phi", "g", "attn"):
                spectral_weight(state, f"{p}self_attn1.snconv1x1_{nm}")
    h = F.relu(batch_norm_train(state, p + "bn", h, momentum=1e-4))              # sagan.py:184
    h = F.conv2d(h, spectral_weight(state, p + "toRGB"), state[p + "toRGB.bias"], 1, 1)
    return torch.tanh(h)


# --------------------------------------------------------------------------------------
# Model forward and the step
# --------------------------------------------------------------------------------------
def get_posterior(state, spec: CelebaSpec, x: Tensor):
    """celeba/module/model.py:157-165."""
    feat = resnet18_features(state, x[..., :3].permute(0, 3, 1, 2))
    h = feat @ state["encoder.fc.weight"].t() + state["encoder.fc.bias"]
    h1, h2 = h[:, : 2 * spec.node], h[:, 2 * spec.node:]
    return h1[:, : spec.node], h1[:, spec.node:], h2[:, : spec.latent_dim], h2[:, spec.latent_dim:]


def transform(state, spec: CelebaSpec, A: Tensor, eps: Tensor):
    u = eps @ A                                                                  # model.py:168
    cols = torch.split(u, 1, dim=1)
    if spec.scm == "linear":
        lat = [base.flow_linear(state[f"flows.{i}.p"], c) for i, c in enumerate(cols)]
    elif spec.scm == "nonlinear":
        lat = [base.flow_planar(state, i, spec.flow_num, c) for i, c in enumerate(cols)]
    else:
        raise ValueError("Not supported SCM!")
    return u.clone(), lat


def decode(state, spec: CelebaSpec, latent: List[Tensor], eps2: Tensor, masks: List[Tensor]):
    """celeba/module/model.py:188-200."""
    zs = [torch.cat([latent[i] for i in idx], dim=1) for idx in DEC_INPUTS] + [eps2]
    sep = [generator(state, k, z) for k, z in enumerate(zs)]
    xs = [o.permute(0, 2, 3, 1) * m for o, m in zip(sep, masks)]
    return sep, torch.tanh(sum(xs))


def forward(state, spec: CelebaSpec, A: Tensor, x: Tensor, masks, noise1: Tensor, noise2: Tensor):
    """celeba/module/model.py:202-218.  Two encode() calls: the second (deterministic) one re-runs the encoder on
    the same input, which changes nothing but the BatchNorm running statistics (updated twice per forward)."""
    mean1, logvar1, mean2, logvar2 = get_posterior(state, spec, x)
    eps1 = mean1 + torch.exp(logvar1 / 2) * noise1                               # model.py:182-183
    eps2 = mean2 + torch.exp(logvar2 / 2) * noise2                               # model.py:184-185
    orig, latent = transform(state, spec, A, eps1)
    sep, xhat = decode(state, spec, latent, eps2, masks)
    m1b, _, _, _ = get_posterior(state, spec, x)                                 # model.py:212 (deterministic pass)
    _, align_latent = transform(state, spec, A, m1b)
    return dict(mean1=mean1, logvar1=logvar1, epsilon1=eps1, orig_latent=orig, latent=latent, mean2=mean2,
                logvar2=logvar2, epsilon2=eps2, align_latent=align_latent, xhat_separated=sep, xhat=xhat)


def step_losses(state, spec: CelebaSpec, A, x, y, masks, noise1, noise2):
    """celeba/module/train.py:27-66."""
    out = forward(state, spec, A, x, masks, noise1, noise2)
    x_ = x[..., :3] * 2 - 1
    recon = (out["xhat"] - x_).abs().sum(dim=[1, 2, 3]).mean()
    kl1 = base.kl_term(out["mean1"], out["logvar1"], spec.node)
    kl2 = base.kl_term(out["mean2"], out["logvar2"], spec.node)                  # subtracts node, train.py:48
    align = base.align_term(out["align_latent"], y[:, : spec.node])
    active = (out["logvar1"].exp().mean(0) < 0.1).float().sum() + (out["logvar2"].exp().mean(0) < 0.1).float().sum()
    active = active / (spec.node + spec.latent_dim)
    loss = recon + spec.beta * (kl1 + kl2) + spec.lam * align
    return loss, {"loss": loss, "recon": recon, "KL": kl1 + kl2, "alignment": align, "active": active}, out


def new_adam_state(state):
    return {k: {"step": 0, "exp_avg": torch.zeros_like(state[k]), "exp_avg_sq": torch.zeros_like(state[k])}
            for k in trainable_names(state)}


def train_step(state: Dict[str, Tensor], adam, spec: CelebaSpec, A, x, y, masks, noise1, noise2):
    """zero_grad -> forward -> losses -> backward -> Adam over the 12,324 trainable parameters.  `state` is updated
    in place (parameters, BatchNorm running statistics, spectral-norm u / v).  Returns (logs, grads, outputs)."""
    names = trainable_names(state)
    work = dict(state)
    for n in names:
        work[n] = state[n].detach().clone().requires_grad_(True)
    loss, logs, out = step_losses(work, spec, A, x, y, masks, noise1, noise2)
    gl = torch.autograd.grad(loss, [work[n] for n in names])
    grads = dict(zip(names, gl))
    with torch.no_grad():
        for n in names:
            base.adam_update(state[n], grads[n], adam[n], spec)
    logs_f = {k: float(v.detach()) for k, v in logs.items()}
    out = {k: ([t.detach() for t in v] if isinstance(v, list) else v.detach()) for k, v in out.items()}
    return logs_f, grads, out
