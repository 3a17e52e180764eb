"""Generate tests/golden/celeba_*.json by running the UNMODIFIED reference celeba/module/{model,train}.py
(imported from /root/reference/celeba) on synthetic inputs and cross-checking oracle/celeba_oracle.py in the
same run.  Build container only:

    python tests/golden/make_golden_celeba.py

`models.resnet18(pretrained=True)` (celeba/module/model.py:117) needs a download; there is no network and no
cached checkpoint, so torchvision's factory is wrapped to return the randomly initialised network.  Nothing
else of the reference is touched.  torch.randn is replaced for the duration of a reference call so that the two
CPU-side draws of model.py:182-185 return pre-generated noise.
"""
import json
import os
import sys

import torch
import torchvision

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from oracle import celeba_oracle as corc  # noqa: E402
from oracle import cdgvae_oracle as orc  # noqa: E402

REF = os.environ.get("CDG_REFERENCE", "/root/reference")
OUT = os.path.dirname(os.path.abspath(__file__))

_orig_resnet18 = torchvision.models.resnet18
torchvision.models.resnet18 = lambda pretrained=False, **kw: _orig_resnet18(weights=None)
sys.path.insert(0, f"{REF}/celeba")
import module.model as rm  # noqa: E402
import module.train as rt  # noqa: E402
rt.tqdm.tqdm = lambda it, **kw: it


def summary(t, nsamp=24):
    t = t.detach().to(torch.float32).reshape(-1)
    n = t.numel()
    idx = torch.unique(torch.linspace(0, n - 1, min(n, nsamp)).round().long())
    d = t.double()
    return {"n": n, "l2": float(d.norm()), "sum": float(d.sum()), "absmax": float(d.abs().max()),
            "idx": idx.tolist(), "val": [float(v) for v in t[idx]]}


def rel(a, b):
    a, b = a.double(), b.double()
    return float((a - b).norm() / (b.norm() + 1e-30))


class NoiseInjector:
    def __init__(self, queue):
        self.queue, self.orig = list(queue), torch.randn

    def __enter__(self):
        def fake(*shape, **kw):
            n = self.queue.pop(0)
            assert tuple(shape) == tuple(n.shape), (shape, n.shape)
            return n.clone()
        torch.randn = fake
        return self

    def __exit__(self, *a):
        torch.randn = self.orig


def full_state(model):
    sd = {k: v.detach().clone() for k, v in model.state_dict().items()}
    for k, g in enumerate(model.decoder):
        for n, v in g.state_dict().items():
            sd[f"decoder.{k}.{n}"] = v.detach().clone()
    return sd


def celeba_case(name, scm, batch, nsteps):
    config = dict(node=6, latent_dim=6, scm=scm, flow_num=1, inverse_loop=100, beta=0.1, cuda=False, lr=1e-3,
                  seed=1, batch_size=batch)
    config["lambda"] = 5.0
    Bm = corc.celeba_B()
    batches = []
    for s in range(nsteps):
        x, y, n1, n2 = corc.synth_celeba(batch, 1234 + s, 4321 + s)
        batches.append(dict(x=x, y=y, noise1=n1, noise2=n2))
    masks = torch.split(batches[0]["x"][..., 3:], 1, dim=-1)           # celeba/main.py:111: first batch's masks
    torch.manual_seed(config["seed"])
    model = rm.CDGVAE(Bm, masks, config, "cpu")
    model.train()
    opt = torch.optim.Adam(model.parameters(), lr=config["lr"])
    spec = corc.CelebaSpec(config)
    A = orc.i_b_inv(Bm)
    assert torch.equal(A, model.I_B_inv)
    state = corc.init_state(config, config["seed"])
    ref0 = full_state(model)
    assert set(state) == set(ref0), (set(state) ^ set(ref0))
    for k in state:
        assert torch.equal(state[k], ref0[k]), ("init", k)
    adam = corc.new_adam_state(state)
    trainable = sorted(n for n, p in model.named_parameters() if p.requires_grad)
    assert trainable == sorted(corc.trainable_names(state))
    # only a few big tensors are summarised in the file; everything is cross-checked here
    keep = lambda k: (k in trainable or "running" in k or "weight_u" in k or "weight_v" in k or k.endswith("num_batches_tracked")
                      or k in ("encoder.conv1.weight", "decoder.0.block0.snlinear0.weight_orig", "decoder.3.block3.conv_1.weight_orig",
                               "decoder.4.toRGB.weight_orig"))
    case = {"name": name, "family": "celeba", "config": {k: v for k, v in config.items() if isinstance(v, (int, float, str, bool))},
            "B": Bm.tolist(), "I_B_inv": A.tolist(), "trainable": trainable,
            "init": {k: summary(v) for k, v in state.items() if keep(k) and "running" not in k and "num_batches" not in k},
            "steps": []}
    worst = 0.0
    for s, b in enumerate(batches, 1):
        entry = {"inputs": {k: summary(v) for k, v in b.items()}}
        with NoiseInjector([b["noise1"], b["noise2"]]):
            logs, xhat = rt.train_CDGVAE([(b["x"], b["y"])], model, config, opt, "cpu")
        entry["logs"] = {k: float(v[0]) for k, v in logs.items()}
        ref_grads = {k: p.grad.detach().clone() for k, p in model.named_parameters() if p.grad is not None}
        assert sorted(ref_grads) == trainable
        ologs, ograds, oout = corc.train_step(state, adam, spec, A, b["x"], b["y"], masks, b["noise1"], b["noise2"])
        for k, v in entry["logs"].items():
            d = abs(ologs[k] - v) / (abs(v) + 1e-12)
            worst = max(worst, d)
            assert d < 1e-5, (name, s, k, ologs[k], v)
        assert rel(oout["xhat"], xhat) < 1e-5
        for k, g in ref_grads.items():
            # fp32 noise floor of this gradient: the oracle at 1 vs 8 threads, and fp32 vs fp64, differ by 1e-3 .. 5e-3
            # relative (L1 loss + train-mode BatchNorm at tiny batch: heavy cancellation); see DESIGN.md §2
            assert rel(ograds[k], g) < 2e-2, (name, s, "grad", k, rel(ograds[k], g))
        ref_state = full_state(model)
        for k in state:
            if k in trainable:
                # Adam's first updates are lr * g / (|g| + eps): elements whose gradient sits at the fp32 noise floor
                # can take the opposite sign, so updated parameters agree to within Adam's 2 * lr bound per step
                assert float((state[k] - ref_state[k]).abs().max()) <= 2 * config["lr"] * s + 1e-7, (name, s, "param", k)
            elif state[k].dtype.is_floating_point:
                assert rel(state[k], ref_state[k]) < 1e-4, (name, s, "state", k, rel(state[k], ref_state[k]))
            else:
                assert torch.equal(state[k], ref_state[k]), (name, s, k)
        # every step is checked one step from identical state: re-synchronise the oracle to the reference
        for k in state:
            state[k].copy_(ref_state[k])
        for n, p in model.named_parameters():
            if p.requires_grad:
                st = opt.state[p]
                adam[n]["exp_avg"].copy_(st["exp_avg"]); adam[n]["exp_avg_sq"].copy_(st["exp_avg_sq"])
                adam[n]["step"] = int(st["step"])
        entry["xhat"] = summary(xhat)
        entry["grads"] = {k: summary(g) for k, g in ref_grads.items()}
        entry["state"] = {k: summary(v) for k, v in ref_state.items() if keep(k)}
        case["steps"].append(entry)
    print(f"{name}: {nsteps} steps, oracle-vs-reference worst log rel diff {worst:.2e}")
    return case


def celeba_free_case(name, batch, nsteps):
    """`nsteps` consecutive reference steps (free-running: nothing is re-synchronised); per-step logs and the trainable
    parameters after the last step.  The product is compared with this trajectory in tests/test_free_running_gpu.py."""
    config = dict(node=6, latent_dim=6, scm="linear", flow_num=1, inverse_loop=100, beta=0.1, cuda=False, lr=1e-3,
                  seed=1, batch_size=batch)
    config["lambda"] = 5.0
    x0 = corc.synth_celeba(batch, 1234, 4321)[0]
    masks = torch.split(x0[..., 3:], 1, dim=-1)
    torch.manual_seed(config["seed"])
    model = rm.CDGVAE(corc.celeba_B(), masks, config, "cpu")
    model.train()
    opt = torch.optim.Adam(model.parameters(), lr=config["lr"])
    case = {"name": name, "family": "celeba", "config": {k: v for k, v in config.items() if isinstance(v, (int, float, str, bool))},
            "steps": []}
    for s in range(nsteps):
        x, y, n1, n2 = corc.synth_celeba(batch, 1234 + s, 4321 + s)
        with NoiseInjector([n1, n2]):
            logs, _ = rt.train_CDGVAE([(x, y)], model, config, opt, "cpu")
        case["steps"].append({"logs": {k: float(v[0]) for k, v in logs.items()}})
    case["final_params"] = {n: summary(p) for n, p in model.named_parameters() if p.requires_grad}
    print(f"{name}: {nsteps} free-running reference steps")
    return case


def main():
    torch.set_num_threads(os.cpu_count())
    if sys.argv[1:] == ["celeba_free6"]:
        c = celeba_free_case("celeba_free6", 2, 6)
        with open(os.path.join(OUT, c["name"] + ".json"), "w") as f:
            json.dump(c, f)
        print("wrote", c["name"])
        return
    # celeba_b16_linear: the reference's own batch (celeba/main.py:70), one step
    cases = {"celeba_linear": ("linear", 2, 2), "celeba_nonlinear": ("nonlinear", 2, 1), "celeba_b16_linear": ("linear", 16, 1)}
    want = sys.argv[1:] or list(cases)
    for c in (celeba_case(n, *cases[n]) for n in want):
        with open(os.path.join(OUT, c["name"] + ".json"), "w") as f:
            json.dump(c, f)
        print("wrote", c["name"], os.path.getsize(os.path.join(OUT, c["name"] + ".json")), "bytes")


if __name__ == "__main__":
    main()
