"""Shared test helpers: golden-summary comparison and case construction (oracle side)."""
import torch

from oracle import cdgvae_oracle as orc


def summary_check(t, g, rtol, name="", atol_scale=1e-7):
    """Compare tensor `t` with a golden summary `g` (see tests/golden/make_golden.py::summary)."""
    t = t.detach().cpu().to(torch.float32).reshape(-1)
    assert t.numel() == g["n"], (name, t.numel(), g["n"])
    d = t.double()
    scale = g["absmax"] + 1e-30
    l2 = float(d.norm())
    assert abs(l2 - g["l2"]) <= rtol * g["l2"] + atol_scale * scale, (name, "l2", l2, g["l2"])
    val = t[torch.tensor(g["idx"])].double()
    ref = torch.tensor(g["val"], dtype=torch.float64)
    err = float((val - ref).abs().max())
    assert err <= rtol * scale + 1e-30, (name, "samples", err, scale)
    return err / scale


def exact_check(t, g, name=""):
    t = t.detach().cpu().to(torch.float32).reshape(-1)
    assert t.numel() == g["n"], name
    val = t[torch.tensor(g["idx"])]
    ref = torch.tensor(g["val"], dtype=torch.float32)
    assert torch.equal(val, ref), (name, val, ref)
    assert float(t.double().sum()) == g["sum"], (name, "sum")


def case_setup(c):
    """Rebuild (spec, B, batches) for a golden case without the reference."""
    cfg = dict(c["config"])
    fam = c["family"]
    name = c["name"]
    nsteps = len(c["steps"])
    batches = []
    if fam == "pendulum":
        mask = orc.pendulum_masks(cfg["image_size"], tuple(c["bands"]))
        spec = orc.pendulum_spec(cfg, mask)
        Bm = orc.pendulum_B(4)
        for s in range(nsteps):
            x, y, noise = orc.synth_pendulum(cfg["batch_size"], cfg["image_size"], 4, 1234 + s, 4321 + s)
            b = dict(x=x, y=y, noise=noise)
            if c["semi"]:
                xl, yl, _ = orc.synth_pendulum(cfg["batch_sizeL"], cfg["image_size"], 4, 9234 + s, 1)
                b.update(x_l=xl, y_l=yl, y=None)
            batches.append(b)
    elif fam == "dr":
        mask = orc.pendulum_masks(cfg["image_size"], tuple(c["bands"]))
        spec = orc.dr_spec(cfg, mask)
        Bm = torch.tensor(c["B"])
        for s in range(nsteps):
            x, y, noise = orc.synth_pendulum(cfg["batch_size"], cfg["image_size"], 5, 1234 + s, 4321 + s)
            batches.append(dict(x=x, y=y, noise=noise))
    elif fam == "vae":
        spec = orc.vae_spec(cfg)
        Bm = orc.pendulum_B(4)
        for s in range(nsteps):
            x, y, noise = orc.synth_pendulum(cfg["batch_size"], cfg["image_size"], 4, 1234 + s, 4321 + s)
            batches.append(dict(x=x, y=y, noise=noise))
    elif fam == "tabular_vae":
        spec = orc.tabular_vae_spec(cfg, c["flatten_topology"])
        Bm = orc.tabular_B(cfg["dataset"])
        for s in range(nsteps):
            x, y, noise = orc.synth_tabular(cfg["dataset"], cfg["batch_size"], 1234 + s, 4321 + s)
            batches.append(dict(x=x, y=y, noise=noise))
    elif fam == "tabular":
        spec = orc.tabular_spec(cfg, c["mask"], c["flatten_topology"])
        Bm = orc.tabular_B(cfg["dataset"])
        for s in range(nsteps):
            x, y, noise = orc.synth_tabular(cfg["dataset"], cfg["batch_size"], 1234 + s, 4321 + s)
            batches.append(dict(x=x, y=y, noise=noise))
    else:
        oil, mask, d, Bm, D = orc.tvae_shape(cfg["dataset"])
        assert mask == c["mask"] and [[list(s) for s in col] for col in oil] == c["output_info_list"]
        spec = orc.tvae_spec(cfg, mask, oil)
        for s in range(nsteps):
            x, y, noise = orc.synth_tvae(cfg["dataset"], cfg["batch_size"], 1234 + s, 4321 + s)
            batches.append(dict(x=x, y=y, noise=noise))
    return spec, Bm, batches, cfg


ALL_CASES = ["pendulum_small_linear", "pendulum_small_nonlinear", "pendulum_small_semi",
             "pendulum_full_linear", "pendulum_full_semi", "tabular_loan", "tabular_adult",
             "tabular_covtype", "tvae_loan", "tvae_covtype", "vae_small_linear", "vae_small_nonlinear", "dr_small_linear",
             "tabular_vae_loan", "tabular_vae_adult", "tabular_vae_covtype"]
