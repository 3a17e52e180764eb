"""The oracle restatement against the committed reference outputs (tests/golden/*.json).
CPU only; needs neither /root/reference nor a GPU."""
import pytest
import torch

from oracle import cdgvae_oracle as orc
from helpers import ALL_CASES, case_setup, exact_check, summary_check


@pytest.mark.parametrize("name", ALL_CASES)
def test_oracle_matches_reference_golden(golden, name):
    c = golden(name)
    spec, Bm, batches, cfg = case_setup(c)
    A = orc.i_b_inv(Bm)
    assert A.tolist() == c["I_B_inv"]                       # bit-exact DAG bookkeeping
    params = orc.init_params(spec, cfg["seed"])
    assert set(params) == set(c["init"])
    for k, v in params.items():                             # same-seed init is bit-exact
        exact_check(v, c["init"][k], k)
    adam = orc.new_adam_state(params)
    for s, (b, e) in enumerate(zip(batches, c["steps"]), 1):
        for k, g in e["inputs"].items():                    # synthetic inputs are bit-exact
            exact_check(b[k], g, k)
        logs, grads, out = orc.train_step(params, adam, spec, A, b["x"], b.get("y"), b["noise"],
                                          b.get("x_l"), b.get("y_l"))
        tol = 2e-6 if s == 1 else 1e-4
        for k, v in e["logs"].items():
            assert abs(logs[k] - v) <= tol * abs(v) + 1e-12, (name, s, k, logs[k], v)
        if "forward" in e:
            f = e["forward"]
            for k in ("mean", "logvar", "epsilon", "orig_latent", "xhat"):    # (VAE: same keys, one decoder)
                summary_check(out[k], f[k], 1e-5, k)
            summary_check(torch.cat(out["latent"], 1), f["latent"], 1e-5, "latent")
            summary_check(torch.cat(out["align_latent"], 1), f["align_latent"], 1e-5, "align_latent")
        if "grads" in e:
            assert sorted(k for k, g in grads.items() if g is None) == e["grad_none"]
            for k, g in e["grads"].items():
                summary_check(grads[k], g, 2e-5, "grad " + k)
        if "params" in e:
            for k, g in e["params"].items():
                summary_check(params[k], g, 1e-5 if s == 1 else 2e-4, f"param {k} step {s}")


def test_flow_inverse_roundtrip():
    """PlanarFlows.inverse (modules/model.py:77-85) undoes forward to fixed-point accuracy."""
    torch.manual_seed(0)
    p = {"flows.0.w.0": torch.randn(1, 1) * 0.1, "flows.0.b.0": torch.randn(1, 1) * 0.1,
         "flows.0.u.0": torch.randn(1, 1) * 0.1}
    h = torch.randn(64, 1)
    z = orc.flow_planar(p, 0, 1, h)
    back = orc.planar_inverse(p, 0, 1, 100, z)
    assert torch.allclose(back, h, atol=1e-5)
