"""CPU oracle for the CDG-VAE / CDG-TVAE training step.  TEST INFRASTRUCTURE ONLY.

This file is a from-scratch *restatement* (torch CPU tensors + autograd, fp32 or
fp64) of the algorithm the reference implements in

    modules/model.py            (CDGVAE, InvertiblePriorLinear, PlanarFlows)
    modules/train.py            (train_CDGVAE, train_CDGVAE_semi)
    tabular/modules/model.py    (CDGVAE, TVAE)
    tabular/modules/train.py    (train_CDGVAE, train_TVAE)
    torch.optim.Adam            (the optimizer every entry script constructs)

It is the *checker* for the CUDA path: only tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs may import it.  The product
package (cdg-vae_b200/) never imports anything from oracle/.

Parity pinning: the reference ships no tests or golden vectors, so
the oracle is pinned against outputs of the *reference itself*, imported unmodified
from /root/reference in the build container by tests/golden/make_golden.py, which
writes tests/golden/*.json.  tests/test_oracle_golden.py re-checks the oracle
against those committed vectors on every run (no /root/reference needed).

Every function cites the reference file:line it restates.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple

import torch
import torch.nn.functional as F

Tensor = torch.Tensor


# --------------------------------------------------------------------------------------
# Spec: what network / loss a config describes
# --------------------------------------------------------------------------------------
@dataclass
class Spec:
    family: str                      # 'pendulum' | 'tabular' | 'tvae'
    node: int
    factor: List[int]
    scm: str                         # 'linear' | 'nonlinear'
    flow_num: int
    input_dim: int                   # flattened width of x
    enc_idx: List[int]               # Sequential indices of the encoder Linears
    dec_idx: List[int]               # Sequential indices of each decoder's Linears
    act: str                         # 'elu' | 'relu'
    n_dec_used: int                  # decoders that take part in decode()
    dataset: str = ""
    image_size: int = 0
    mask: Optional[list] = None      # pendulum: list of (H,W,3) tensors; tabular: list of int
    flatten_topology: Optional[List[int]] = None
    output_info_list: Optional[list] = None   # tvae: list[list[(dim, activation_fn)]]
    sigma_range: Optional[Tuple[float, float]] = None
    beta: float = 1.0
    lam: float = 1.0
    lr: float = 1e-3
    weight_decay: float = 0.0
    betas: Tuple[float, float] = (0.9, 0.999)
    eps: float = 1e-8
    single_decoder: bool = False      # the VAE baseline (modules/model.py:102-189): one decoder over all latents
    dr: bool = False                  # DR variant (DR/modules/model.py:245, :284-287): decoders also see the last latent


def pendulum_spec(config: dict, mask: Sequence[Tensor]) -> Spec:
    """modules/model.py:209-250 (CDGVAE.__init__) + main.py:93-107 (defaults)."""
    assert sum(config["factor"]) == config["node"]          # model.py:214
    assert len(config["factor"]) == len(mask)               # model.py:215
    return Spec(
        family="pendulum", node=config["node"], factor=list(config["factor"]),
        scm=config["scm"], flow_num=config.get("flow_num", 1),
        input_dim=3 * config["image_size"] ** 2, enc_idx=[0, 2, 4], dec_idx=[0, 2, 4],
        act="elu", n_dec_used=len(mask), image_size=config["image_size"], mask=list(mask),
        beta=config["beta"], lam=config["lambda"], lr=config.get("lr", 1e-3),
    )


def dr_spec(config: dict, mask: Sequence[Tensor]) -> Spec:
    """DR/modules/model.py:209-250: like CDGVAE but `sum(factor) == node - 1` and every decoder's first Linear takes
    its factor's latents plus the last ("spurious") latent: nn.Linear(k+1, 300)."""
    assert len(config["factor"]) == len(mask)
    assert sum(config["factor"]) == config["node"] - 1
    sp = pendulum_spec(dict(config, node=sum(config["factor"])), mask)
    sp.node = config["node"]
    sp.dr = True
    return sp


def vae_spec(config: dict) -> Spec:
    """modules/model.py:102-140 (VAE.__init__): single decoder Linear(node,300)-ELU-Linear-ELU-Linear-Tanh."""
    s = config["image_size"]
    return Spec(
        family="pendulum", node=config["node"], factor=[config["node"]], scm=config["scm"],
        flow_num=config.get("flow_num", 1), input_dim=3 * s * s, enc_idx=[0, 2, 4], dec_idx=[0, 2, 4], act="elu",
        n_dec_used=1, image_size=s, mask=[torch.ones(s, s, 3)], beta=config["beta"], lam=config["lambda"],
        lr=config.get("lr", 1e-3), single_decoder=True,
    )


def tabular_spec(config: dict, mask: Sequence[int], flatten_topology: Optional[Sequence[int]]) -> Spec:
    """tabular/modules/model.py:234-305 (CDGVAE.__init__)."""
    assert sum(config["factor"]) == config["node"]
    assert len(config["factor"]) == len(mask)
    cov = config["dataset"] == "covtype"
    return Spec(
        family="tabular", node=config["node"], factor=list(config["factor"]),
        scm=config["scm"], flow_num=config.get("flow_num", 1), input_dim=config["input_dim"],
        enc_idx=[0, 2, 4, 6] if cov else [0, 2], dec_idx=[0, 2, 4] if cov else [0, 2],
        act="elu", n_dec_used=len(mask), dataset=config["dataset"], mask=list(mask),
        flatten_topology=None if flatten_topology is None else list(flatten_topology),
        beta=config["beta"], lam=config["lambda"], lr=config.get("lr", 1e-2),
    )


def tabular_vae_spec(config: dict, flatten_topology: Optional[Sequence[int]]) -> Spec:
    """tabular/modules/model.py:103-172 (VAE.__init__): the CDG-VAE encoder, ONE decoder over all latents
    (loan node-4-D, adult / covtype node-8-8-16-D')."""
    ds = config["dataset"]
    cov = ds == "covtype"
    out = config["input_dim"] - 1 + 7 if cov else config["input_dim"]
    return Spec(
        family="tabular", node=config["node"], factor=[config["node"]], scm=config["scm"], flow_num=config.get("flow_num", 1),
        input_dim=config["input_dim"], enc_idx=[0, 2, 4, 6] if cov else [0, 2], dec_idx=[0, 2] if ds == "loan" else [0, 2, 4, 6],
        act="elu", n_dec_used=1, dataset=ds, mask=[out],
        flatten_topology=None if flatten_topology is None else list(flatten_topology),
        beta=config["beta"], lam=config["lambda"], lr=config.get("lr", 1e-2), single_decoder=True,
    )


def tvae_spec(config: dict, mask: Sequence[int], output_info_list) -> Spec:
    """tabular/modules/model.py:360-407 (TVAE.__init__); tabular/main_tvae.py:196-200."""
    assert sum(config["factor"]) == config["node"]
    assert len(config["factor"]) == len(mask)
    oil = [[(int(s[0]), str(s[1])) for s in col] for col in output_info_list]
    return Spec(
        family="tvae", node=config["node"], factor=list(config["factor"]),
        scm=config["scm"], flow_num=config.get("flow_num", 1), input_dim=config["input_dim"],
        enc_idx=[0, 2, 4, 6], dec_idx=[0, 2, 4, 6], act="relu", n_dec_used=len(mask),
        dataset=config.get("dataset", ""), mask=list(mask), output_info_list=oil,
        sigma_range=tuple(config["sigma_range"]), beta=1.0, lam=config["lambda"],
        lr=config.get("lr", 1e-3), weight_decay=config.get("weight_decay", 0.0),
    )


# --------------------------------------------------------------------------------------
# Model pieces
# --------------------------------------------------------------------------------------
def _act(h: Tensor, kind: str) -> Tensor:
    return F.elu(h) if kind == "elu" else F.relu(h)


def mlp(params: Dict[str, Tensor], prefix: str, idx: Sequence[int], h: Tensor, act: str) -> Tensor:
    """nn.Sequential(Linear, act, Linear, act, ..., Linear): model.py:219-225, :243-250."""
    for n, i in enumerate(idx):
        h = h @ params[f"{prefix}.{i}.weight"].t() + params[f"{prefix}.{i}.bias"]
        if n + 1 < len(idx):
            h = _act(h, act)
    return h


def get_posterior(params, spec: Spec, x: Tensor) -> Tuple[Tensor, Tensor]:
    """modules/model.py:256-259; tabular/modules/model.py:311-314."""
    h = mlp(params, "encoder", spec.enc_idx, x.reshape(x.shape[0], -1), spec.act)
    return h[:, : spec.node], h[:, spec.node:]


def i_b_inv(B: Tensor) -> Tensor:
    """modules/model.py:228-230: I_B_inv = inverse(I - B)."""
    return torch.inverse(torch.eye(B.shape[0], dtype=B.dtype) - B)


def flow_linear(p: Tensor, u: Tensor) -> Tensor:
    """InvertiblePriorLinear.forward, modules/model.py:20-25."""
    return p[0] * u + p[1]


def planar_build_u(u_: Tensor, w_: Tensor) -> Tensor:
    """PlanarFlows.build_u, modules/model.py:70-75 (log(1+exp(.)) written literally)."""
    wu = w_.t() @ u_
    term1 = -1 + torch.log(1 + torch.exp(wu))
    return u_ + (term1 - wu) * (w_ / torch.norm(w_, p=2) ** 2)


def flow_planar(params, i: int, flow_num: int, h: Tensor) -> Tensor:
    """PlanarFlows.forward, modules/model.py:87-100 (input_dim = 1)."""
    for j in range(flow_num):
        w, b, u = params[f"flows.{i}.w.{j}"], params[f"flows.{i}.b.{j}"], params[f"flows.{i}.u.{j}"]
        u_hat = planar_build_u(u, w)
        h = h + u_hat.t() * F.elu(h @ w + b)
    return h


def planar_inverse(params, i: int, flow_num: int, inverse_loop: int, h: Tensor) -> Tensor:
    """PlanarFlows.inverse, modules/model.py:77-85 (fixed-point iteration)."""
    for j in reversed(range(flow_num)):
        w, b, u = params[f"flows.{i}.w.{j}"], params[f"flows.{i}.b.{j}"], params[f"flows.{i}.u.{j}"]
        z = h
        for _ in range(inverse_loop):
            z = h - planar_build_u(u, w).t() * F.elu(z @ w + b)
        h = z
    return h


def transform(params, spec: Spec, A: Tensor, eps: Tensor) -> Tuple[Tensor, List[Tensor]]:
    """CDGVAE.transform, modules/model.py:261-268: u = eps @ I_B_inv, then one flow per node."""
    u = eps @ A
    orig = u.clone()
    cols = torch.split(u, 1, dim=1)
    if spec.scm == "linear":
        latent = [flow_linear(params[f"flows.{i}.p"], c) for i, c in enumerate(cols)]
    elif spec.scm == "nonlinear":
        latent = [flow_planar(params, i, spec.flow_num, c) for i, c in enumerate(cols)]
    else:
        raise ValueError("Not supported SCM!")          # model.py:240
    return orig, latent


def decode(params, spec: Spec, latent: List[Tensor]) -> Tuple[List[Tensor], Tensor]:
    """CDGVAE.decode: modules/model.py:281-288 (mask, sum, tanh);
    tabular/modules/model.py:337-342 and :439-444 (plain cat)."""
    zc = torch.cat(latent, dim=1)
    if spec.dr:                                                 # DR/modules/model.py:284-287
        spurious = zc[:, [-1]]
        z = [torch.cat([t, spurious], dim=1) for t in torch.split(zc[:, :-1], spec.factor, dim=-1)]
    else:
        z = torch.split(zc, spec.factor, dim=-1)
    if spec.single_decoder:                                     # VAE.forward, modules/model.py:178-180
        o = mlp(params, "decoder", spec.dec_idx, z[0], spec.act)
        if spec.family != "pendulum":                           # tabular VAE: xhat = decoder(cat(latent)), tabular/modules/model.py:222
            return [o], o
        return [o], torch.tanh(o).view(-1, spec.image_size, spec.image_size, 3)
    sep = [mlp(params, f"decoder.{k}", spec.dec_idx, z[k], spec.act) for k in range(spec.n_dec_used)]
    if spec.family == "pendulum":
        s = spec.image_size
        xs = [o.view(-1, s, s, 3) * m.to(o.dtype) for o, m in zip(sep, spec.mask)]
        return sep, torch.tanh(sum(xs))
    return sep, torch.cat(sep, dim=1)


def forward(params, spec: Spec, A: Tensor, x: Tensor, noise: Tensor):
    """CDGVAE.forward, modules/model.py:290-304.  The encoder is evaluated once: the
    reference's second (deterministic) pass recomputes the same `mean` (model.py:300)."""
    mean, logvar = get_posterior(params, spec, x)
    eps = mean + torch.exp(logvar / 2) * noise                 # model.py:277
    orig, latent = transform(params, spec, A, eps)
    sep, xhat = decode(params, spec, latent)
    _, align_latent = transform(params, spec, A, mean)         # deterministic: eps = mean
    return dict(mean=mean, logvar=logvar, epsilon=eps, orig_latent=orig, latent=latent,
                align_latent=align_latent, xhat_separated=sep, xhat=xhat)


# --------------------------------------------------------------------------------------
# Losses
# --------------------------------------------------------------------------------------
def kl_term(mean: Tensor, logvar: Tensor, node: int) -> Tensor:
    """modules/train.py:180-185."""
    kl = mean.pow(2).sum(1) - logvar.sum(1) + logvar.exp().sum(1) - node
    return (0.5 * kl).mean()


def align_term(align_latent: List[Tensor], y: Tensor) -> Tensor:
    """modules/train.py:189-190: BCE on sigmoid probabilities (log clamped at -100)."""
    y_hat = torch.sigmoid(torch.cat(align_latent, dim=1))
    return F.binary_cross_entropy(y_hat, y, reduction="none").sum(1).mean()


def recon_term(params, spec: Spec, xhat: Tensor, x: Tensor) -> Tensor:
    if spec.family == "pendulum":                               # modules/train.py:175
        return 0.5 * (xhat - x).pow(2).sum(dim=[1, 2, 3]).mean()
    if spec.family == "tabular":                                # tabular/modules/train.py:199-210
        if spec.dataset == "loan":
            return 0.5 * (xhat - x[:, spec.flatten_topology]).pow(2).sum(1).mean()
        if spec.dataset == "adult":
            x_ = x[:, spec.flatten_topology]
            r = 0.5 * (xhat[:, :2] - x_[:, :2]).pow(2).sum(1).mean()
            r = r + 0.5 * (xhat[:, 3:] - x_[:, 3:]).pow(2).sum(1).mean()
            return r + F.binary_cross_entropy_with_logits(xhat[:, [2]], x_[:, [2]])
        if spec.dataset == "covtype":
            r = 0.5 * (xhat[:, :7] - x[:, :7]).pow(2).sum(1).mean()
            return r + F.nll_loss(F.log_softmax(xhat[:, 7:], 1), (x[:, 7] - 1).to(torch.int64))
        raise ValueError("Not supported dataset!")              # train.py:210
    if spec.family == "tvae":                                   # tabular/modules/train.py:270-285
        sigma = params["sigma"]
        start, recon = 0, 0
        for column_info in spec.output_info_list:
            for dim, fn in column_info:
                end = start + dim
                if fn != "softmax":
                    std = sigma[start]
                    residual = x[:, start] - torch.tanh(xhat[:, start])
                    recon = recon + (residual ** 2 / 2 / (std ** 2)).mean() + torch.log(std)
                else:
                    recon = recon + F.cross_entropy(
                        xhat[:, start:end], torch.argmax(x[:, start:end], dim=-1), reduction="mean")
                start = end
        return recon
    raise ValueError(spec.family)


def step_losses(params, spec: Spec, A: Tensor, x: Tensor, y: Tensor, noise: Tensor,
                x_l: Optional[Tensor] = None, y_l: Optional[Tensor] = None):
    """One batch of train_CDGVAE (modules/train.py:170-199), train_CDGVAE_semi (:241-275),
    tabular train_CDGVAE (tabular/modules/train.py:192-236) or train_TVAE (:264-308)."""
    out = forward(params, spec, A, x, noise)
    recon = recon_term(params, spec, out["xhat"], x)
    kl = kl_term(out["mean"], out["logvar"], spec.node)
    if x_l is not None:                                         # semi: modules/train.py:261-263
        mean_l, _ = get_posterior(params, spec, x_l)
        _, al = transform(params, spec, A, mean_l)
        align = align_term(al, y_l[:, : spec.node])
    elif spec.family == "pendulum":
        align = align_term(out["align_latent"], y[:, : spec.node])
    else:                                                       # tabular: all of y (train.py:224)
        align = align_term(out["align_latent"], y)
    var_ = out["logvar"].exp().mean(0)
    loss = recon + spec.beta * kl + spec.lam * align
    logs = {"loss": loss, "recon": recon, "KL": kl, "alignment": align}
    for i in range(spec.node):
        logs[f"posterior_variance{i + 1}"] = var_[i]
    return loss, logs, out


# --------------------------------------------------------------------------------------
# Adam (torch.optim.Adam single-tensor path, amsgrad=False, maximize=False)
# --------------------------------------------------------------------------------------
def adam_update(p: Tensor, g: Tensor, st: dict, spec: Spec) -> None:
    """torch/optim/adam.py::_single_tensor_adam as used via main.py:189-192,
    tabular/main_tvae.py:196-200 (coupled weight decay g += wd*p)."""
    b1, b2 = spec.betas
    st["step"] += 1
    t = st["step"]
    if spec.weight_decay != 0:
        g = g + spec.weight_decay * p
    st["exp_avg"].lerp_(g, 1 - b1)
    st["exp_avg_sq"].mul_(b2).addcmul_(g, g, value=1 - b2)
    bc1 = 1 - b1 ** t
    bc2 = 1 - b2 ** t
    step_size = spec.lr / bc1
    denom = (st["exp_avg_sq"].sqrt() / math.sqrt(bc2)).add_(spec.eps)
    p.addcdiv_(st["exp_avg"], denom, value=-step_size)


def new_adam_state(params: Dict[str, Tensor]) -> Dict[str, dict]:
    return {k: {"step": 0, "exp_avg": torch.zeros_like(v), "exp_avg_sq": torch.zeros_like(v)}
            for k, v in params.items()}


def train_step(params: Dict[str, Tensor], adam: Dict[str, dict], spec: Spec, A: Tensor,
               x: Tensor, y: Tensor, noise: Tensor,
               x_l: Optional[Tensor] = None, y_l: Optional[Tensor] = None):
    """zero_grad -> forward -> losses -> backward -> Adam (-> sigma clamp for TVAE).
    `params` are leaf tensors updated in place.  Returns (logs as floats, grads, outputs)."""
    leaves = {k: v.detach().requires_grad_(True) for k, v in params.items()}
    loss, logs, out = step_losses(leaves, spec, A, x, y, noise, x_l, y_l)
    names = list(leaves)
    gl = torch.autograd.grad(loss, [leaves[n] for n in names], allow_unused=True)
    grads = dict(zip(names, gl))
    with torch.no_grad():
        for n in names:
            if grads[n] is None:                    # Adam skips params without a grad
                continue
            adam_update(params[n], grads[n], adam[n], spec)
        if spec.family == "tvae":                   # tabular/modules/train.py:314
            params["sigma"].clamp_(spec.sigma_range[0], spec.sigma_range[1])
    logs_f = {k: float(v.detach()) for k, v in logs.items()}
    out = {k: ([t.detach() for t in v] if isinstance(v, list) else v.detach()) for k, v in out.items()}
    return logs_f, grads, out


# --------------------------------------------------------------------------------------
# InfoMax baseline (main.py --model InfoMax): VAE + a discriminator on (x, epsilon)
# --------------------------------------------------------------------------------------
def init_discriminator(config: dict, hidden: int = 300) -> Dict[str, Tensor]:
    """Discriminator.__init__, modules/model.py:191-201: Linear(P + node, 300)-ELU-Linear(300, 300)-ELU-Linear(300, 1),
    created right after the VAE under the same RNG stream (main.py:155-157); nn.Linear supplies the init law."""
    import torch.nn as nn
    dims = [3 * config["image_size"] ** 2 + config["node"], hidden, hidden, 1]
    sd = {}
    for j, (i, o) in zip((0, 2, 4), zip(dims[:-1], dims[1:])):
        lin = nn.Linear(i, o)
        sd[f"net.{j}.weight"], sd[f"net.{j}.bias"] = lin.weight.detach().clone(), lin.bias.detach().clone()
    return sd


def discriminator(dparams, x: Tensor, z: Tensor) -> Tensor:
    """Discriminator.forward, modules/model.py:203-206."""
    h = torch.cat((x.reshape(x.shape[0], -1), z), dim=1)
    return mlp(dparams, "net", [0, 2, 4], h, "elu")


def infomax_train_step(params, dparams, adam, adam_d, spec: Spec, A: Tensor, x: Tensor, y: Tensor, noise: Tensor, perm: Tensor,
                       gamma: float, lr_d: float):
    """One batch of train_InfoMax, modules/train.py:71-148.  `loss.backward(retain_graph=True); MI.backward()` accumulates
    d loss + d MI into every .grad (train.py:139-140), i.e. the gradient of recon + beta KL + lambda align + (gamma + 1) MI;
    both optimizers then step (train.py:141-142)."""
    leaves = {k: v.detach().requires_grad_(True) for k, v in params.items()}
    dleaves = {k: v.detach().requires_grad_(True) for k, v in dparams.items()}
    out = forward(leaves, spec, A, x, noise)
    recon = recon_term(leaves, spec, out["xhat"], x)
    kl = kl_term(out["mean"], out["logvar"], spec.node)
    align = align_term(out["align_latent"], y[:, : spec.node])
    d_joint = discriminator(dleaves, x, out["epsilon"])
    d_marg = discriminator(dleaves, x, out["epsilon"][perm])            # permute_dims, train.py:73-77
    mi = -(d_joint.mean() - torch.exp(d_marg - 1).mean())
    loss = recon + spec.beta * kl + spec.lam * align + gamma * mi
    names, dnames = list(leaves), list(dleaves)
    gl = torch.autograd.grad(loss + mi, [leaves[n] for n in names] + [dleaves[n] for n in dnames])
    grads = dict(zip(names, gl[: len(names)]))
    dgrads = dict(zip(dnames, gl[len(names):]))
    spec_d = Spec(**{**spec.__dict__, "lr": lr_d})
    with torch.no_grad():
        for n in names:
            adam_update(params[n], grads[n], adam[n], spec)
        for n in dnames:
            adam_update(dparams[n], dgrads[n], adam_d[n], spec_d)
    logs = {"loss": loss, "recon": recon, "KL": kl, "alignment": align, "MutualInfo": mi}
    var_ = out["logvar"].exp().mean(0)
    for i in range(spec.node):
        logs[f"posterior_variance{i + 1}"] = var_[i]
    return {k: float(v.detach()) for k, v in logs.items()}, grads, dgrads, {k: (v.detach() if torch.is_tensor(v) else v) for k, v in out.items()}


# --------------------------------------------------------------------------------------
# Synthetic inputs of SURVEY.md §8(d): defined in synthetic_inputs.py (shared with bench.py and smoke(),
# which must not depend on oracle/ for their inputs); re-exported here for the tests and the golden scripts
# --------------------------------------------------------------------------------------
import os as _os
import sys as _sys

_sys.path.insert(0, _os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))))
from synthetic_inputs import (pendulum_B, pendulum_masks, tabular_B, synth_pendulum, synth_tabular, tvae_shape,  # noqa: E402,F401
                              synth_tvae)


# --------------------------------------------------------------------------------------
# Same-seed initialisation (parameter creation order of the reference constructors)
# --------------------------------------------------------------------------------------
def init_params(spec: Spec, seed: int = 1, hidden: int = 300) -> Dict[str, Tensor]:
    """Create a state dict exactly as `torch.manual_seed(seed); Model(B, mask, config, 'cpu')`
    would: encoder Linears -> flows -> decoders (-> sigma).  modules/model.py:219-250,
    tabular/modules/model.py:245-305 and :371-407.  nn.Linear supplies the init law."""
    import torch.nn as nn
    torch.manual_seed(seed)
    sd: Dict[str, Tensor] = {}

    def seq(prefix, dims, idx):
        for (i, o), j in zip(zip(dims[:-1], dims[1:]), idx):
            lin = nn.Linear(i, o)
            sd[f"{prefix}.{j}.weight"] = lin.weight.detach().clone()
            sd[f"{prefix}.{j}.bias"] = lin.bias.detach().clone()

    d, fam = spec.node, spec.family
    if fam == "pendulum":
        enc = [spec.input_dim, hidden, hidden, 2 * d]
        decs = [[k + (1 if spec.dr else 0), hidden, hidden, spec.input_dim] for k in spec.factor]
        extra = []
    elif fam == "tabular" and spec.single_decoder:                      # tabular VAE, model.py:110-172
        enc = [spec.input_dim, 4, 4, 4, 2 * d] if spec.dataset == "covtype" else [spec.input_dim, 4, 2 * d]
        decs = [[d, 4, spec.mask[0]]] if spec.dataset == "loan" else [[d, 8, 8, 16, spec.mask[0]]]
        extra = []
    elif fam == "tabular":
        if spec.dataset == "covtype":
            enc = [spec.input_dim, 4, 4, 4, 2 * d]
            decs = [[k, 2, 2, m] for k, m in zip(spec.factor, spec.mask)]
            extra = [[spec.factor[-1], 4, 4, 8, spec.mask[-1]]]       # the unused 7th decoder
        else:
            enc = [spec.input_dim, 4, 2 * d]
            decs = [[k, 2, m] for k, m in zip(spec.factor, spec.mask)]
            extra = []
    else:
        enc = [spec.input_dim, 32, 16, 16, 2 * d]
        decs = [[k, 8, 8, 16, m] for k, m in zip(spec.factor, spec.mask)]
        extra = []
    seq("encoder", enc, range(0, 2 * len(enc), 2))
    for i in range(d):
        if spec.scm == "linear":
            sd[f"flows.{i}.p"] = torch.rand([2]) * 0.1                 # model.py:18
        else:
            for nm in ("w", "b", "u"):                                 # model.py:60-68
                for j in range(spec.flow_num):
                    sd[f"flows.{i}.{nm}.{j}"] = torch.randn(1, 1) * 0.1
    for k, dims in enumerate(decs + extra):
        seq("decoder" if spec.single_decoder else f"decoder.{k}", dims, range(0, 2 * len(dims), 2))
    if fam == "tvae":
        sd["sigma"] = torch.ones(spec.input_dim) * 0.1                 # model.py:407
    return sd
