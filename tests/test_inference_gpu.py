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


def _ref_planar(w, b, u, h, inverse_loop=None, log_determinant=False):
    """modules/model.py:70-100 restated in fp64 (input_dim = 1): build_u, forward with log|det|, fixed-point inverse."""
    F = len(w)
    uh = [u[j] + ((-1 + torch.log(1 + torch.exp(w[j] * u[j]))) - w[j] * u[j]) * (w[j] / w[j].abs() ** 2) for j in range(F)]
    elu = torch.nn.functional.elu
    if inverse_loop is not None:
        for j in reversed(range(F)):
            z = h
            for _ in range(inverse_loop):
                z = h - uh[j] * elu(z * w[j] + b[j])
            h = z
        return h
    logdet = torch.zeros_like(h)
    for j in range(F):
        x = h * w[j] + b[j]
        grad = torch.where(x > 0, torch.ones_like(x), torch.exp(x))
        logdet = logdet + torch.log((1 + grad * w[j] * uh[j]).abs())
        h = h + uh[j] * elu(x)
    return (h, logdet) if log_determinant else h


@pytest.mark.parametrize("scm,flow_num", [("linear", 1), ("nonlinear", 1), ("nonlinear", 2)])
def test_flow_kernel_forward_logdet_inverse(scm, flow_num):
    """cdg_flow_apply behind model.inverse / model.transform(log_determinant=True) / flows[i](x) / flows[i].inverse(x):
    one launch instead of the reference's 100 x ~6 eager launches per node (modules/model.py:77-85)."""
    from cdgvae_b200.modules.model import CDGVAE
    cfg = dict(node=4, scm=scm, flow_num=flow_num, inverse_loop=100, factor=[1, 1, 2], image_size=8, lr=1e-3, beta=0.1, seed=1)
    cfg["lambda"] = 5.0
    Bm = orc.pendulum_B(4)
    torch.manual_seed(3)
    model = CDGVAE(Bm, orc.pendulum_masks(8, (3, 6)), cfg, "cpu")
    with torch.no_grad():                                     # larger flow parameters than the 0.1-scale init: a real test of the inverse
        for p in model.flows.parameters():
            p.mul_(6.0)
    model = model.to("cuda")
    g = torch.Generator().manual_seed(5)
    eps = torch.randn(257, 4, generator=g)
    u_ref = eps.double() @ orc.i_b_inv(Bm).double()
    sd = {k: v.detach().double().cpu() for k, v in model.state_dict().items()}
    z_ref, ld_ref, inv_ref = [], [], []
    for i in range(4):
        h = u_ref[:, i:i + 1]
        if scm == "linear":
            p = sd[f"flows.{i}.p"]
            z, ld = p[0] * h + p[1], torch.log(p[0].abs()).repeat(h.shape[0], 1)        # model.py:20-25
            inv = (h - p[1]) / p[0]                                                      # model.py:27-29
        else:
            w = [sd[f"flows.{i}.w.{j}"].reshape(()) for j in range(flow_num)]
            b = [sd[f"flows.{i}.b.{j}"].reshape(()) for j in range(flow_num)]
            u = [sd[f"flows.{i}.u.{j}"].reshape(()) for j in range(flow_num)]
            z, ld = _ref_planar(w, b, u, h, log_determinant=True)
            inv = _ref_planar(w, b, u, h, inverse_loop=100)
        z_ref.append(z); ld_ref.append(ld); inv_ref.append(inv)
    orig, latent, logdet = model.transform(eps.cuda(), log_determinant=True)
    assert rel(orig, u_ref) < 1e-5
    assert isinstance(latent, list) and latent[0].shape == (257, 1) and logdet[0].shape == (257, 1)
    assert rel(torch.cat(latent, 1), torch.cat(z_ref, 1)) < 1e-5
    assert float((torch.cat(logdet, 1).double().cpu() - torch.cat(ld_ref, 1)).abs().max()) < 2e-6
    _, _, nold = model.transform(eps.cuda())
    assert nold == [0] * 4                                                                # model.py:22, :89
    cols = list(torch.split(orig, 1, dim=1))
    inv = model.inverse(cols)                                                             # model.py:252-254
    assert rel(torch.cat(inv, 1), torch.cat(inv_ref, 1)) < 1e-5
    # the per-node modules, called the way inference.py:317 / :308 do
    for i, layer in enumerate(model.flows):
        o, ld = layer(cols[i], log_determinant=True)
        assert rel(o, z_ref[i]) < 1e-5 and float((ld.double().cpu() - ld_ref[i]).abs().max()) < 2e-6
        assert layer(cols[i])[1] == 0
        assert rel(layer.inverse(cols[i]), inv_ref[i]) < 1e-5
    # round trip: inverse(flow(u)) = u
    back = model.inverse(latent)
    assert rel(torch.cat(back, 1), orig) < 1e-4
    # forward(log_determinant=True) returns the same log|det| list (model.py:290-304)
    x, _, nz = orc.synth_pendulum(5, 8, 4, 1, 2)
    model.noise_fn = lambda n, d: nz
    out = model(x.cuda(), log_determinant=True)
    _, _, ld2 = model.transform(out[2], log_determinant=True)
    assert rel(torch.cat(out[5], 1), torch.cat(ld2, 1)) < 1e-6


def test_flow_modules_have_no_cpu_path():
    from cdgvae_b200.modules.model import InvertiblePriorLinear, PlanarFlows
    with pytest.raises(RuntimeError):
        InvertiblePriorLinear()(torch.zeros(3, 1))
    with pytest.raises(RuntimeError):
        PlanarFlows(1, 1, 100).inverse(torch.zeros(3, 1))
