"""Generate tests/golden/*.json by running the UNMODIFIED reference (imported by file path
from /root/reference) on the synthetic inputs of SURVEY.md §8(d), and cross-check the oracle
restatement (oracle/cdgvae_oracle.py) against it in the same run.

Run in the build container only (it needs /root/reference):

    python tests/golden/make_golden.py

The reference has no tests or golden vectors of its own; these files are the
pin.  Each file stores, per case: the config, seeds, checksums of the synthetic inputs,
summaries (shape, l2, sum, absmax, strided samples) of the initial parameters, per-step
logs, forward outputs of step 1, gradients of step 1, parameters after selected steps and
the Adam state after the last step.
"""
import importlib.util
import json
import os
import sys
from collections import namedtuple

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import cdgvae_oracle as orc  # noqa: E402

REF = os.environ.get("CDG_REFERENCE", "/root/reference")
OUT = os.path.dirname(os.path.abspath(__file__))


def load(name, path):
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


pm = load("ref_pend_model", f"{REF}/modules/model.py")
pt = load("ref_pend_train", f"{REF}/modules/train.py")
tm = load("ref_tab_model", f"{REF}/tabular/modules/model.py")
tt = load("ref_tab_train", f"{REF}/tabular/modules/train.py")
for m in (pt, tt):
    m.tqdm.tqdm = lambda it, **kw: it            # silence progress bars


def summary(t, nsamp=24):
    t = t.detach().to(torch.float32).reshape(-1)
    n = t.numel()
    idx = torch.unique(torch.linspace(0, n - 1, min(n, nsamp)).round().long())
    d = t.double()
    return {"n": n, "l2": float(d.norm()), "sum": float(d.sum()), "absmax": float(d.abs().max()),
            "idx": idx.tolist(), "val": [float(v) for v in t[idx]]}


def summarize_dict(d):
    return {k: summary(v) for k, v in d.items() if v is not None}


class NoiseInjector:
    """Replace torch.randn for the duration of a reference call so that the CPU-side draw at
    modules/model.py:276 returns pre-generated noise."""

    def __init__(self, queue):
        self.queue = list(queue)
        self.orig = torch.randn

    def __enter__(self):
        def fake(*shape, **kw):
            n = self.queue.pop(0)
            assert tuple(shape) == tuple(n.shape), (shape, n.shape)
            return n.clone()
        torch.randn = fake
        return self

    def __exit__(self, *a):
        torch.randn = self.orig


def rel(a, b):
    a, b = a.double(), b.double()
    return float((a - b).norm() / (b.norm() + 1e-30))


def run_case(name, family, config, model, spec, Bmat, batches, ref_step, steps_keep=(1, 2)):
    """batches: list of dict(x,y,noise[,x_l,y_l]).  ref_step(batch) runs ONE reference train step
    and returns its logs dict (lists of length 1) [and maybe xhat]."""
    A = orc.i_b_inv(Bmat)
    assert torch.equal(A, model.I_B_inv), "I_B_inv restatement differs"
    params = {k: v.detach().clone() for k, v in model.state_dict().items()}
    adam = orc.new_adam_state(params)
    case = {"name": name, "family": family,
            "config": {k: v for k, v in config.items() if isinstance(v, (int, float, str, list, bool))},
            "I_B_inv": A.tolist(), "init": summarize_dict(params), "steps": []}
    nsteps = len(batches)
    worst = 0.0
    for s, b in enumerate(batches, 1):
        entry = {"inputs": summarize_dict({k: v for k, v in b.items() if v is not None})}
        # ---- reference ----
        if s == 1:
            with NoiseInjector([b["noise"]]):
                out = model(b["x"])
            names = ["mean", "logvar", "epsilon", "orig_latent", "latent", "logdet", "align_latent",
                     "xhat_separated", "xhat"]
            if len(out) == 8:                       # the VAE baseline has no xhat_separated (model.py:189)
                names.remove("xhat_separated")
            fo = dict(zip(names, out))
            entry["forward"] = summarize_dict({
                "mean": fo["mean"], "logvar": fo["logvar"], "epsilon": fo["epsilon"],
                "orig_latent": fo["orig_latent"], "latent": torch.cat(fo["latent"], 1),
                "align_latent": torch.cat(fo["align_latent"], 1), "xhat": fo["xhat"]})
        with NoiseInjector([b["noise"]]):
            logs = ref_step(b)
        entry["logs"] = {k: float(v[0]) for k, v in logs.items()}
        ref_grads = {k: p.grad.detach().clone() for k, p in model.named_parameters() if p.grad is not None}
        # ---- oracle on the same inputs ----
        ologs, ograds, _ = orc.train_step(params, adam, spec, A, b["x"], b.get("y"), b["noise"],
                                          b.get("x_l"), b.get("y_l"))
        for k in entry["logs"]:
            d = abs(ologs[k] - entry["logs"][k]) / (abs(entry["logs"][k]) + 1e-12)
            worst = max(worst, d)
            assert d < (2e-6 if s == 1 else 1e-4), (name, s, k, ologs[k], entry["logs"][k])
        for k, g in ref_grads.items():
            d = rel(ograds[k], g)
            assert d < (1e-5 if s == 1 else 1e-3), (name, s, "grad", k, d)
        assert set(k for k, g in ograds.items() if g is not None) == set(ref_grads), "grad None-ness differs"
        sd = model.state_dict()
        for k in params:
            d = rel(params[k], sd[k])
            assert d < (1e-5 if s == 1 else 1e-3), (name, s, "param", k, d)
        if s == 1:
            entry["grads"] = summarize_dict(ref_grads)
            entry["grad_none"] = sorted(k for k, p in model.named_parameters() if p.grad is None)
        if s in steps_keep or s == nsteps:
            entry["params"] = summarize_dict({k: v for k, v in sd.items()})
        case["steps"].append(entry)
    print(f"{name}: {nsteps} steps, oracle-vs-reference worst log rel diff {worst:.2e}")
    return case


# ------------------------------------------------------------------------------------------
def pendulum_case(name, scm, image_size, bands, batch, nsteps, semi=False, batch_l=0):
    config = dict(node=4, scm=scm, flow_num=1, inverse_loop=100, factor=[1, 1, 2], image_size=image_size,
                  batch_size=batch, batch_sizeL=batch_l, lr=1e-3, beta=0.1, cuda=False, seed=1)
    config["lambda"] = 5.0
    Bm = orc.pendulum_B(4)
    mask = orc.pendulum_masks(image_size, bands)
    torch.manual_seed(config["seed"])
    model = pm.CDGVAE(Bm, mask, config, "cpu")
    model.train()
    opt = torch.optim.Adam(model.parameters(), lr=config["lr"])
    batches = []
    for s in range(nsteps):
        x, y, noise = orc.synth_pendulum(batch, image_size, 4, seed=1234 + s, noise_seed=4321 + s)
        b = dict(x=x, y=y, noise=noise)
        if semi:
            xl, yl, _ = orc.synth_pendulum(batch_l, image_size, 4, seed=9234 + s, noise_seed=1)
            b.update(x_l=xl, y_l=yl, y=None)
        batches.append(b)
    spec = orc.pendulum_spec(config, mask)

    if not semi:
        def ref_step(b):
            logs, _ = pt.train_CDGVAE([(b["x"], b["y"])], model, config, opt, "cpu")
            return logs
    else:
        class FakeLoader:
            def __init__(self, ds, batch_size, shuffle):
                self.ds = ds
            def __iter__(self):
                return iter(self.ds)
        pt.DataLoader = FakeLoader

        def ref_step(b):
            logs, _ = pt.train_CDGVAE_semi([(b["x_l"], b["y_l"])], [b["x"]], model, config, opt, "cpu")
            return logs
    case = run_case(name, "pendulum", config, model, spec, Bm, batches, ref_step)
    case["bands"] = list(bands)
    case["semi"] = semi
    return case


def vae_case(name, scm, image_size, batch, nsteps):
    config = dict(node=4, scm=scm, flow_num=1, inverse_loop=100, image_size=image_size, batch_size=batch, lr=1e-3,
                  beta=0.1, cuda=False, seed=1)
    config["lambda"] = 5.0
    Bm = orc.pendulum_B(4)
    torch.manual_seed(config["seed"])
    model = pm.VAE(Bm, config, "cpu")
    model.train()
    opt = torch.optim.Adam(model.parameters(), lr=config["lr"])
    batches = []
    for s in range(nsteps):
        x, y, noise = orc.synth_pendulum(batch, image_size, 4, seed=1234 + s, noise_seed=4321 + s)
        batches.append(dict(x=x, y=y, noise=noise))
    spec = orc.vae_spec(config)

    def ref_step(b):
        logs, _ = pt.train_VAE([(b["x"], b["y"])], model, config, opt, "cpu")
        return logs
    case = run_case(name, "vae", config, model, spec, Bm, batches, ref_step)
    case["semi"] = False
    return case


def dr_case(name, scm, image_size, bands, batch, nsteps):
    drm = load("ref_dr_model", f"{REF}/DR/modules/model.py")
    config = dict(node=5, scm=scm, flow_num=1, inverse_loop=100, factor=[1, 1, 2], image_size=image_size,
                  batch_size=batch, lr=1e-3, beta=0.1, cuda=False, seed=1)
    config["lambda"] = 5.0
    Bm = torch.zeros(5, 5)
    Bm[0, 2] = Bm[0, 3] = Bm[1, 2] = Bm[1, 3] = 1
    indeg = Bm.sum(0); m = indeg != 0
    Bm[:, m] = Bm[:, m] / indeg[m]                      # DR/main.py:137-147
    mask = orc.pendulum_masks(image_size, bands)
    torch.manual_seed(config["seed"])
    model = drm.CDGVAE(Bm, mask, config, "cpu")
    model.train()
    opt = torch.optim.Adam(model.parameters(), lr=config["lr"])
    batches = []
    for s in range(nsteps):
        x, y, noise = orc.synth_pendulum(batch, image_size, 5, seed=1234 + s, noise_seed=4321 + s)
        batches.append(dict(x=x, y=y, noise=noise))
    spec = orc.dr_spec(config, mask)

    def ref_step(b):
        logs, _ = pt.train_CDGVAE([(b["x"], b["y"])], model, config, opt, "cpu")
        return logs
    case = run_case(name, "dr", config, model, spec, Bm, batches, ref_step)
    case["bands"] = list(bands)
    case["semi"] = False
    case["B"] = Bm.tolist()
    return case


def infomax_case(name, image_size, batch, nsteps):
    """main.py --model InfoMax: VAE + Discriminator, train_InfoMax (modules/train.py:71-148).  torch.randperm is replaced
    for the duration of the reference call so that permute_dims uses a pre-generated permutation."""
    config = dict(node=4, scm="linear", flow_num=1, inverse_loop=100, image_size=image_size, batch_size=batch, lr=1e-3,
                  lr_D=1e-4, beta=0.1, gamma=1.0, cuda=False, seed=1)
    config["lambda"] = 5.0
    Bm = orc.pendulum_B(4)
    torch.manual_seed(config["seed"])
    model = pm.VAE(Bm, config, "cpu")
    disc = pm.Discriminator(config, "cpu")
    model.train(); disc.train()
    opt = torch.optim.Adam(model.parameters(), lr=config["lr"])
    opt_d = torch.optim.Adam(disc.parameters(), lr=config["lr_D"])
    spec = orc.vae_spec(config)
    A = orc.i_b_inv(Bm)
    torch.manual_seed(config["seed"])
    params = orc.init_params(spec, config["seed"])
    dparams = orc.init_discriminator(config)
    for k, v in model.state_dict().items():
        assert torch.equal(v, params[k]), k
    for k, v in disc.state_dict().items():
        assert torch.equal(v, dparams[k]), k
    adam, adam_d = orc.new_adam_state(params), orc.new_adam_state(dparams)
    case = {"name": name, "family": "infomax", "config": {k: v for k, v in config.items() if isinstance(v, (int, float, str, bool))},
            "I_B_inv": A.tolist(), "init": summarize_dict(params), "init_d": summarize_dict(dparams), "steps": []}
    worst = 0.0
    for s in range(nsteps):
        x, y, noise = orc.synth_pendulum(batch, image_size, 4, seed=1234 + s, noise_seed=4321 + s)
        perm = torch.randperm(batch, generator=torch.Generator().manual_seed(99 + s))
        orig_perm = torch.randperm
        torch.randperm = lambda n, **kw: perm.clone()
        try:
            with NoiseInjector([noise]):
                logs, xhat = pt.train_InfoMax([(x, y)], model, disc, config, opt, opt_d, "cpu")
        finally:
            torch.randperm = orig_perm
        entry = {"perm": perm.tolist(), "logs": {k: float(v[0]) for k, v in logs.items()}}
        ol, og, odg, _ = orc.infomax_train_step(params, dparams, adam, adam_d, spec, A, x, y, noise, perm, config["gamma"], config["lr_D"])
        for k, v in entry["logs"].items():
            d = abs(ol[k] - v) / (abs(v) + 1e-12)
            worst = max(worst, d)
            assert d < (2e-6 if s == 0 else 1e-4), (name, s, k, ol[k], v)
        for n, p in model.named_parameters():
            assert rel(og[n], p.grad) < (1e-5 if s == 0 else 1e-3), (name, s, n)
        for n, p in disc.named_parameters():
            assert rel(odg[n], p.grad) < (1e-5 if s == 0 else 1e-3), (name, s, "D", n)
        if s == 0:
            entry["grads"] = summarize_dict({n: p.grad for n, p in model.named_parameters()})
            entry["grads_d"] = summarize_dict({n: p.grad for n, p in disc.named_parameters()})
        entry["params_d"] = summarize_dict(disc.state_dict())
        case["steps"].append(entry)
    print(f"{name}: {nsteps} steps, oracle-vs-reference worst log rel diff {worst:.2e}")
    return case


def tabular_case(dataset, batch, nsteps):
    config = dict(dataset=dataset, scm="linear", flow_num=1, inverse_loop=100, batch_size=batch, lr=0.01,
                  beta=0.01, cuda=False, seed=1)
    config["lambda"] = 10.0
    if dataset in ("loan", "adult"):
        config.update(node=3, factor=[1, 1, 1], input_dim=5)
        mask = [2, 2, 1] if dataset == "loan" else [1, 1, 3]
        ft = [1, 2, 3, 4, 0] if dataset == "loan" else [2, 3, 0, 1, 4]
    else:
        config.update(node=6, factor=[1] * 6, input_dim=8)
        mask = [1, 1, 2, 1, 1, 8]
        ft = None
    Bm = orc.tabular_B(dataset)
    torch.manual_seed(config["seed"])
    model = tm.CDGVAE(Bm, mask, config, "cpu")
    model.train()
    opt = torch.optim.Adam(model.parameters(), lr=config["lr"])
    DS = namedtuple("DS", ["flatten_topology"])
    ds = DS(ft)
    batches = []
    for s in range(nsteps):
        x, y, noise = orc.synth_tabular(dataset, batch, seed=1234 + s, noise_seed=4321 + s)
        batches.append(dict(x=x, y=y, noise=noise))
    spec = orc.tabular_spec(config, mask, ft)

    def ref_step(b):
        return tt.train_CDGVAE(ds, [(b["x"], b["y"])], model, config, opt, "cpu")
    case = run_case(f"tabular_{dataset}", "tabular", config, model, spec, Bm, batches, ref_step)
    case["mask"] = mask
    case["flatten_topology"] = ft
    return case


def tabular_vae_case(dataset, batch, nsteps):
    """The single-decoder baseline of the tabular tree: tabular/modules/model.py::VAE + tabular/modules/train.py::train_VAE."""
    config = dict(dataset=dataset, scm="linear", flow_num=1, inverse_loop=100, batch_size=batch, lr=0.01,
                  beta=0.01, cuda=False, seed=1)
    config["lambda"] = 10.0
    if dataset in ("loan", "adult"):
        config.update(node=3, input_dim=5)
        ft = [1, 2, 3, 4, 0] if dataset == "loan" else [2, 3, 0, 1, 4]
    else:
        config.update(node=6, input_dim=8)
        ft = None
    Bm = orc.tabular_B(dataset)
    torch.manual_seed(config["seed"])
    model = tm.VAE(Bm, config, "cpu")
    model.train()
    opt = torch.optim.Adam(model.parameters(), lr=config["lr"])
    DS = namedtuple("DS", ["flatten_topology"])
    ds = DS(ft)
    batches = []
    for s in range(nsteps):
        x, y, noise = orc.synth_tabular(dataset, batch, seed=1234 + s, noise_seed=4321 + s)
        batches.append(dict(x=x, y=y, noise=noise))
    spec = orc.tabular_vae_spec(config, ft)

    def ref_step(b):
        return tt.train_VAE(ds, [(b["x"], b["y"])], model, config, opt, "cpu")
    case = run_case(f"tabular_vae_{dataset}", "tabular_vae", config, model, spec, Bm, batches, ref_step)
    case["mask"] = spec.mask
    case["flatten_topology"] = ft
    return case


def tvae_case(kind, batch, nsteps):
    oil, mask, d, Bm, D = orc.tvae_shape(kind)
    sr = [0.01, 0.1] if kind == "loan" else [0.005, 0.01]
    config = dict(dataset=kind, scm="linear", flow_num=1, inverse_loop=100, batch_size=batch, lr=1e-3,
                  weight_decay=1e-5, cuda=False, seed=1, node=d, factor=[1] * d, input_dim=D, sigma_range=sr)
    config["lambda"] = 5.0
    torch.manual_seed(config["seed"])
    model = tm.TVAE(Bm, mask, config, "cpu")
    model.train()
    opt = torch.optim.Adam(model.parameters(), lr=config["lr"], weight_decay=config["weight_decay"])
    SpanInfo = namedtuple("SpanInfo", ["dim", "activation_fn"])
    ref_oil = [[SpanInfo(*s) for s in col] for col in oil]
    batches = []
    for s in range(nsteps):
        x, y, noise = orc.synth_tvae(kind, batch, seed=1234 + s, noise_seed=4321 + s)
        batches.append(dict(x=x, y=y, noise=noise))
    spec = orc.tvae_spec(config, mask, oil)

    def ref_step(b):
        return tt.train_TVAE(ref_oil, None, [(b["x"], b["y"])], model, config, opt, "cpu")
    case = run_case(f"tvae_{kind}", "tvae", config, model, spec, Bm, batches, ref_step)
    case["mask"] = mask
    case["output_info_list"] = oil
    return case


def main():
    torch.set_num_threads(os.cpu_count())
    if len(sys.argv) > 1 and sys.argv[1] == "infomax":
        c = infomax_case("infomax_small", 8, 16, 3)
        with open(os.path.join(OUT, c["name"] + ".json"), "w") as f:
            json.dump(c, f)
        print("wrote", c["name"])
        return
    if len(sys.argv) > 1 and sys.argv[1] == "dr":
        c = dr_case("dr_small_linear", "linear", 8, (3, 6), 16, 4)
        with open(os.path.join(OUT, c["name"] + ".json"), "w") as f:
            json.dump(c, f)
        print("wrote", c["name"])
        return
    if len(sys.argv) > 1 and sys.argv[1] == "tabular_vae":  # round 2: the tabular tree's baseline
        for c in (tabular_vae_case(d, 256, 4) for d in ("loan", "adult", "covtype")):
            with open(os.path.join(OUT, c["name"] + ".json"), "w") as f:
                json.dump(c, f)
            print("wrote", c["name"])
        return
    if len(sys.argv) > 1 and sys.argv[1] == "vae":         # added after the first batch of goldens
        cases = [vae_case("vae_small_linear", "linear", 8, 16, 4), vae_case("vae_small_nonlinear", "nonlinear", 8, 16, 4)]
        for c in cases:
            with open(os.path.join(OUT, c["name"] + ".json"), "w") as f:
                json.dump(c, f)
            print("wrote", c["name"])
        return
    cases = [
        pendulum_case("pendulum_small_linear", "linear", 8, (3, 6), 16, 4),
        pendulum_case("pendulum_small_nonlinear", "nonlinear", 8, (3, 6), 16, 4),
        pendulum_case("pendulum_small_semi", "nonlinear", 8, (3, 6), 16, 4, semi=True, batch_l=4),
        pendulum_case("pendulum_full_linear", "linear", 64, (20, 51), 128, 6),
        pendulum_case("pendulum_full_semi", "nonlinear", 64, (20, 51), 128, 4, semi=True, batch_l=32),
        tabular_case("loan", 256, 6),
        tabular_case("adult", 256, 6),
        tabular_case("covtype", 256, 6),
        tvae_case("loan", 256, 6),
        tvae_case("covtype", 256, 6),
    ]
    for c in cases:
        with open(os.path.join(OUT, c["name"] + ".json"), "w") as f:
            json.dump(c, f)
        print("wrote", c["name"], os.path.getsize(os.path.join(OUT, c["name"] + ".json")), "bytes")


if __name__ == "__main__":
    main()
