"""Free-running trajectories: SIX consecutive training steps of the product from the reference's initial state, compared
with six consecutive steps of the UNMODIFIED reference (tests/golden/*.json were minted that way), nothing re-synchronised
in between.  One case per family.

This is the literal reading of the north-star's "updated parameters after N steps".  The other parity tests check every step
one step from identical state (DESIGN.md §2 explains why: Adam's first updates are lr * g / (|g| + eps), so an element
whose gradient is at rounding-noise level moves by +-lr for ANY implementation, the reference on another BLAS included,
and trajectories separate at that rate).  Here the achieved deviation is asserted against the bound written in DESIGN.md
§2 and recorded (gpurun_out/free_running.json) whenever the test runs on a GPU box."""
import json
import os
from collections import namedtuple

import pytest
import torch

from helpers import case_setup

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

# family -> (golden, bound on |log - golden| / |golden| over the six steps, bound on the final parameters).
# Parameters: per-tensor ||p - p_ref||_2 / ||p_ref||_2 from the golden's (l2, strided samples) summary.
CASES = {
    "pendulum": ("pendulum_full_linear", 5e-4, 3e-3),
    "tabular": ("tabular_adult", 5e-4, 3e-3),
    "tvae": ("tvae_loan", 5e-4, 3e-3),
    # CelebA: the reference's own trajectory moves by up to 1e-1 in the logs over six steps when only its BLAS thread count
    # changes (tools/celeba_free_noise.py -> profiles/r02_celeba_free_running_noise.json: L1 loss through train-mode BatchNorm
    # at batch 2, Adam's sign-like first updates).  Per-step bound = 3 x that recorded drift (never below 2e-3); parameters:
    # Adam's 2 * lr * steps bound.
    "celeba": ("celeba_free6", None, None),
}


def _celeba_log_bounds():
    d = json.load(open(os.path.join(ROOT, "profiles", "r02_celeba_free_running_noise.json")))
    per = [max(v["log_rel_dev_per_step"][s] for v in d.values()) for s in range(6)]
    return [max(2e-3, 3.0 * p) for p in per]


def _record(family, entry):
    out = os.path.join(ROOT, "gpurun_out")
    if not os.path.isdir(out):
        return
    path = os.path.join(out, "free_running.json")
    data = json.load(open(path)) if os.path.exists(path) else {}
    data[family] = entry
    json.dump(data, open(path, "w"), indent=1)


def _sample_dev(t, g):
    """max |sample - golden sample| / absmax, and | l2 - golden l2 | / l2, from a golden summary."""
    t = t.detach().cpu().to(torch.float32).reshape(-1)
    val = t[torch.tensor(g["idx"])].double()
    ref = torch.tensor(g["val"], dtype=torch.float64)
    return float((val - ref).abs().max()) / (g["absmax"] + 1e-30), abs(float(t.double().norm()) - g["l2"]) / (g["l2"] + 1e-30)


@pytest.mark.parametrize("family", list(CASES))
def test_six_free_running_steps_against_the_reference(golden, family):
    name, log_tol, par_tol = CASES[family]
    c = golden(name)
    assert len(c["steps"]) == 6
    if family == "celeba":
        from oracle import celeba_oracle as corc
        from cdgvae_b200.celeba.module.model import CDGVAE
        from cdgvae_b200.celeba.module.train import train_CDGVAE
        cfg = dict(c["config"], pretrained=False)
        x0 = corc.synth_celeba(cfg["batch_size"], 1234, 4321)[0]
        torch.manual_seed(cfg["seed"])
        model = CDGVAE(corc.celeba_B(), torch.split(x0[..., 3:], 1, dim=-1), cfg, "cpu").to("cuda")
        opt = torch.optim.Adam(model.parameters(), lr=cfg["lr"])
        before = {n: p.detach().clone() for n, p in model.named_parameters() if p.requires_grad}

        def step(s):
            x, y, n1, n2 = corc.synth_celeba(cfg["batch_size"], 1234 + s, 4321 + s)
            q = [n1, n2]
            model.noise_fn = lambda b, d: q.pop(0)
            return train_CDGVAE([(x, y)], model, cfg, opt, "cuda")[0]
    else:
        spec, Bm, batches, cfg = case_setup(c)
        torch.manual_seed(cfg["seed"])
        if family == "pendulum":
            from cdgvae_b200.modules.model import CDGVAE
            from cdgvae_b200.modules import train as T
            model = CDGVAE(Bm, spec.mask, cfg, "cpu").to("cuda")
            opt = torch.optim.Adam(model.parameters(), lr=cfg["lr"])
            run = lambda b: T.train_CDGVAE([(b["x"], b["y"])], model, cfg, opt, "cuda")[0]
        else:
            from cdgvae_b200.tabular.modules import model as M, train as T
            if family == "tvae":
                model = M.TVAE(Bm, c["mask"], cfg, "cpu").to("cuda")
                opt = torch.optim.Adam(model.parameters(), lr=cfg["lr"], weight_decay=cfg["weight_decay"])
                Span = namedtuple("SpanInfo", ["dim", "activation_fn"])
                oil = [[Span(*sp) for sp in col] for col in c["output_info_list"]]
                run = lambda b: T.train_TVAE(oil, None, [(b["x"], b["y"])], model, cfg, opt, "cuda")
            else:
                model = M.CDGVAE(Bm, c["mask"], cfg, "cpu").to("cuda")
                opt = torch.optim.Adam(model.parameters(), lr=cfg["lr"])
                DS = namedtuple("DS", ["flatten_topology"])
                run = lambda b: T.train_CDGVAE(DS(c["flatten_topology"]), [(b["x"], b["y"])], model, cfg, opt, "cuda")

        def step(s):
            b = batches[s]
            model.noise_fn = lambda n, d: b["noise"]
            return run(b)

    worst_log, per_step = 0.0, []
    log_bounds = _celeba_log_bounds() if family == "celeba" else [log_tol] * 6
    for s, e in enumerate(c["steps"]):
        logs = step(s)
        dev = max(abs(logs[k][0] - v) / (abs(v) + 1e-12) for k, v in e["logs"].items() if abs(v) > 1e-6)
        per_step.append(dev)
        worst_log = max(worst_log, dev)
    entry = {"golden": name, "steps": 6, "log_rel_dev_per_step": per_step, "log_rel_dev_max": worst_log, "log_bound_per_step": log_bounds}
    sd = dict(model.named_parameters())
    if family == "celeba":
        moved = max(float((sd[n].detach() - before[n]).abs().max()) for n in before)
        devs = {n: _sample_dev(sd[n], g) for n, g in c["final_params"].items()}
        entry.update(param_sample_dev_max=max(d[0] for d in devs.values()), param_l2_dev_max=max(d[1] for d in devs.values()),
                     param_bound="2 * lr * 6 per element")
        _record(family, entry)
        assert moved <= 6 * 1.001 * cfg["lr"] + 1e-7
        for n, g in c["final_params"].items():
            t = sd[n].detach().cpu().reshape(-1)[torch.tensor(g["idx"])].double()
            assert float((t - torch.tensor(g["val"], dtype=torch.float64)).abs().max()) <= 2 * cfg["lr"] * 6 + 1e-7, n
    else:
        final = c["steps"][-1]["params"]
        devs = {n: _sample_dev(model.state_dict()[n], g) for n, g in final.items()}
        entry.update(param_sample_dev_max=max(d[0] for d in devs.values()), param_l2_dev_max=max(d[1] for d in devs.values()),
                     param_bound=par_tol)
        _record(family, entry)
        for n, (ds, dl) in devs.items():
            assert ds <= par_tol and dl <= par_tol, (family, n, ds, dl)
    assert all(d <= b for d, b in zip(per_step, log_bounds)), (family, per_step, log_bounds)
