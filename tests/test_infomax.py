"""InfoMax baseline (main.py --model InfoMax: modules/model.py VAE + Discriminator, modules/train.py:71-148 train_InfoMax).
CPU: the oracle against the committed reference golden.  GPU: the drop-in against the oracle and the golden."""
import pytest
import torch

from oracle import cdgvae_oracle as orc
from helpers import exact_check, summary_check

RTOL = 1e-4


def rel(a, b):
    a, b = a.detach().double().cpu().reshape(-1), b.detach().double().cpu().reshape(-1)
    return float((a - b).norm() / (b.norm() + 1e-30))


def _setup(c):
    cfg = dict(c["config"])
    spec = orc.vae_spec(cfg)
    Bm = orc.pendulum_B(4)
    torch.manual_seed(cfg["seed"])
    params = orc.init_params(spec, cfg["seed"])
    dparams = orc.init_discriminator(cfg)            # created right after the VAE under the same RNG stream (main.py:155-157)
    return cfg, spec, Bm, params, dparams


def _batch(cfg, s):
    x, y, noise = orc.synth_pendulum(cfg["batch_size"], cfg["image_size"], 4, seed=1234 + s, noise_seed=4321 + s)
    return x, y, noise


def test_infomax_oracle_matches_reference_golden(golden):
    c = golden("infomax_small")
    cfg, spec, Bm, params, dparams = _setup(c)
    for k, g in c["init"].items():
        exact_check(params[k], g, k)
    for k, g in c["init_d"].items():
        exact_check(dparams[k], g, "D " + k)
    adam, adam_d = orc.new_adam_state(params), orc.new_adam_state(dparams)
    A = orc.i_b_inv(Bm)
    for s, e in enumerate(c["steps"]):
        x, y, noise = _batch(cfg, s)
        logs, grads, dgrads, _ = orc.infomax_train_step(params, dparams, adam, adam_d, spec, A, x, y, noise, torch.tensor(e["perm"]),
                                                       cfg["gamma"], cfg["lr_D"])
        for k, v in e["logs"].items():
            assert abs(logs[k] - v) <= (2e-6 if s == 0 else 1e-4) * abs(v) + 1e-12, (s, k, logs[k], v)
        if "grads" in e:
            for k, g in e["grads"].items():
                summary_check(grads[k], g, 2e-5, "grad " + k)
            for k, g in e["grads_d"].items():
                summary_check(dgrads[k], g, 2e-5, "D grad " + k)
        for k, g in e["params_d"].items():
            summary_check(dparams[k], g, 1e-5 if s == 0 else 2e-4, f"D param {k} step {s}")


@pytest.mark.gpu
def test_infomax_step_matches_reference_and_oracle(golden):
    from cdgvae_b200.modules.model import VAE, Discriminator
    from cdgvae_b200.modules.train import train_InfoMax
    c = golden("infomax_small")
    cfg, spec, Bm, oparams, odparams = _setup(c)
    torch.manual_seed(cfg["seed"])
    model = VAE(Bm, cfg, "cpu")
    disc = Discriminator(cfg, "cpu")
    for k, v in disc.state_dict().items():
        assert torch.equal(v, odparams[k]), k                      # same-seed construction, same keys
    model, disc = model.to("cuda"), disc.to("cuda")
    opt = torch.optim.Adam(model.parameters(), lr=cfg["lr"])
    opt_d = torch.optim.Adam(disc.parameters(), lr=cfg["lr_D"])
    oadam, oadam_d = orc.new_adam_state(oparams), orc.new_adam_state(odparams)
    A = orc.i_b_inv(Bm)
    for s, e in enumerate(c["steps"]):
        x, y, noise = _batch(cfg, s)
        perm = torch.tensor(e["perm"])
        if s > 0:      # one step from identical state
            for (m, op, oa, o) in ((model, oparams, oadam, opt), (disc, odparams, oadam_d, opt_d)):
                for n, p in m.named_parameters():
                    op[n].copy_(p.detach().cpu())
                    st = o.state[p]
                    oa[n]["exp_avg"].copy_(st["exp_avg"].cpu()); oa[n]["exp_avg_sq"].copy_(st["exp_avg_sq"].cpu())
                    oa[n]["step"] = int(st["step"])
        model.noise_fn = lambda n, d: noise
        model.perm_fn = lambda n: perm
        logs, xhat = train_InfoMax([(x, y)], model, disc, cfg, opt, opt_d, "cuda")
        ol, og, odg, oo = orc.infomax_train_step(oparams, odparams, oadam, oadam_d, spec, A, x, y, noise, perm, cfg["gamma"], cfg["lr_D"])
        assert list(logs) == list(e["logs"])
        for k, v in ol.items():
            assert abs(logs[k][0] - v) <= RTOL * abs(v) + 1e-7, (s, k, logs[k][0], v)
            if s == 0:
                assert abs(logs[k][0] - e["logs"][k]) <= RTOL * abs(e["logs"][k]) + 1e-7, (s, k, "golden")
        assert rel(xhat, oo["xhat"]) < RTOL
        for n, p in model.named_parameters():
            assert rel(p.grad, og[n]) < RTOL, (s, n, rel(p.grad, og[n]))
            if "grads" in e:
                summary_check(p.grad, e["grads"][n], RTOL, "golden grad " + n, atol_scale=1e-6)
        for n, p in disc.named_parameters():
            assert rel(p.grad, odg[n]) < RTOL, (s, "D", n, rel(p.grad, odg[n]))
            if "grads_d" in e:
                summary_check(p.grad, e["grads_d"][n], RTOL, "golden D grad " + n, atol_scale=1e-6)
        for n, p in disc.named_parameters():                        # both optimizers stepped (train.py:141-142)
            assert float((p.detach().cpu() - odparams[n]).abs().max()) <= 2.2 * cfg["lr_D"], n
            assert int(opt_d.state[p]["step"]) == s + 1
