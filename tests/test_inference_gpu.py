"""The evaluation-side use of the drop-in model: the do-intervention loop of the reference's
inference.py:298-327 / metric.py:226-255 — `encode(deterministic=True)`, `model.inverse` (PlanarFlows fixed-point
inverse, modules/model.py:77-85), recomputation of the descendants through `model.B`, the flows, `model.decode` —
run against the drop-in on the GPU and against the oracle on the CPU."""
import pytest
import torch

from oracle import cdgvae_oracle as orc

pytestmark = pytest.mark.gpu


def rel(a, b):
    a, b = a.detach().double().cpu().reshape(-1), b.detach().double().cpu().reshape(-1)
    return float((a - b).norm() / (b.norm() + 1e-30))


def intervene(latent, epsilon, Bmat, do_index, do_value, inverse, flow):
    """inference.py:302-315, written once for both sides."""
    node = len(latent)
    latent_ = [t.clone() for t in latent]
    latent_[do_index] = torch.full_like(latent_[do_index], do_value)
    z = torch.cat(inverse(latent_), dim=1).clone()
    for j in range(node):
        if j == do_index:
            continue
        if j == 0:
            z[:, j] = epsilon[:, j]
        z[:, j] = torch.matmul(z[:, :j], Bmat[:j, j]) + epsilon[:, j]
    return flow(list(torch.split(z, 1, dim=1)))


@pytest.mark.parametrize("scm", ["linear", "nonlinear"])
def test_do_intervention_matches_oracle(scm):
    from cdgvae_b200.modules.model import CDGVAE
    cfg = dict(node=4, scm=scm, flow_num=1, inverse_loop=100, factor=[1, 1, 2], image_size=64, lr=1e-3, beta=0.1, seed=1)
    cfg["lambda"] = 5.0
    Bm, mask = orc.pendulum_B(4), orc.pendulum_masks(64)
    spec = orc.pendulum_spec(cfg, mask)
    torch.manual_seed(1)
    model = CDGVAE(Bm, mask, cfg, "cpu").to("cuda")
    model.B = model.B.cuda()
    params = orc.init_params(spec, 1)
    A = orc.i_b_inv(Bm)
    x, _, _ = orc.synth_pendulum(8, 64, 4, 77, 78)
    # product
    mean, logvar, epsilon, orig_latent, latent, logdet = model.encode(x.cuda(), deterministic=True)
    assert len(latent) == 4 and latent[0].shape == (8, 1) and logdet == [0] * 4
    # oracle
    omean, _ = orc.get_posterior(params, spec, x)
    _, olatent = orc.transform(params, spec, A, omean)
    assert rel(mean, omean) < 1e-4 and rel(torch.cat(latent, 1), torch.cat(olatent, 1)) < 1e-4

    def oinverse(lat):
        if scm == "linear":
            return [(t - params[f"flows.{i}.p"][1]) / params[f"flows.{i}.p"][0] for i, t in enumerate(lat)]
        return [orc.planar_inverse(params, i, 1, 100, t) for i, t in enumerate(lat)]

    def oflow(cols):
        if scm == "linear":
            return [orc.flow_linear(params[f"flows.{i}.p"], c) for i, c in enumerate(cols)]
        return [orc.flow_planar(params, i, 1, c) for i, c in enumerate(cols)]

    with torch.no_grad():
        for do_index in range(4):
            for do_value in (-0.3, 0.4):
                z = intervene(latent, epsilon, model.B, do_index, do_value, model.inverse,
                              lambda cols: [layer(c)[0] for c, layer in zip(cols, model.flows)])
                sep, do_xhat = model.decode(z)
                oz = intervene(olatent, omean, Bm, do_index, do_value, oinverse, oflow)
                osep, oxhat = orc.decode(params, spec, oz)
                assert rel(torch.cat(z, 1), torch.cat(oz, 1)) < 1e-4, (do_index, do_value)
                assert do_xhat.shape == (8, 64, 64, 3) and rel(do_xhat, oxhat) < 1e-4
                for k in range(3):
                    assert rel(sep[k], osep[k]) < 1e-4          # per-factor images shown by inference.py:286-289
