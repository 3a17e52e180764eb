"""CelebA CDG-VAE on the GPU, through the C ABI, against the CPU oracle (oracle/celeba_oracle.py) and the committed
reference goldens (tests/golden/celeba_*.json).

Tolerances.  Forward quantities (logs, xhat, latents, BatchNorm running statistics, spectral-norm vectors): 1e-4
relative.  Gradients of this step sit on an fp32 noise floor of 1e-3 .. 5e-3 relative — measured on the reference's own
arithmetic: its gradients at 1 vs 8 CPU threads, and in fp32 vs fp64, differ by that much (L1 reconstruction through
five train-mode-BatchNorm generators at tiny batch: heavy cancellation).  They are therefore compared with the fp64
oracle at GRAD_RTOL = 1e-2, i.e. the CUDA path must be as close to the exact gradient as the reference itself is
(tools/celeba_grad_noise.py prints the three-way comparison: fp32 oracle, tcgen05 3xTF32 path, SIMT fp32 path); the 12
flow scalars are compared as one gradient vector (a component that is tiny next to its neighbours carries the absolute
noise of the whole vector).  The single layers are compared with torch at 1e-5.  Updated parameters: 1e-4 relative on
every element whose gradient stands clear of the noise floor (adam_param_check below), Adam's 2*lr bound on the rest.
Batches: 2 (two teacher-forced steps) and the reference's own 16 (celeba/main.py:70).
"""
import ctypes as C

import pytest
import torch
import torch.nn.functional as F

from oracle import cdgvae_oracle as orc
from oracle import celeba_oracle as corc
from helpers import summary_check

pytestmark = pytest.mark.gpu
RTOL = 1e-4
GRAD_RTOL = 1e-2


def rel(a, b):
    a, b = a.detach().double().cpu().reshape(-1), b.detach().double().cpu().reshape(-1)
    return float((a - b).norm() / (b.norm() + 1e-30))


def _conv_desc(cin, cout, k, stride, pad):
    from cdgvae_b200 import _lib
    c = _lib.Conv()
    c.w, c.b, c.u, c.v = 0, -1, -1, -1
    c.cin, c.cout, c.k, c.stride, c.pad = cin, cout, k, stride, pad
    return c


@pytest.mark.parametrize("B,H,cin,cout,k,stride,pad,up", [
    (2, 16, 64, 32, 3, 1, 1, 1), (2, 8, 128, 64, 3, 1, 1, 2), (3, 32, 3, 64, 7, 2, 3, 1), (2, 16, 64, 128, 3, 2, 1, 1),
    (2, 16, 64, 128, 1, 2, 0, 1), (2, 8, 32, 3, 3, 1, 1, 1), (1, 4, 512, 512, 3, 1, 1, 2), (2, 8, 256, 128, 1, 1, 0, 1)])
def test_conv2d_forward_matches_torch(B, H, cin, cout, k, stride, pad, up):
    from cdgvae_b200 import _lib
    L = _lib.lib()
    g = torch.Generator().manual_seed(B * 1000 + cin + cout + k)
    x = torch.randn(B, H, H, cin, generator=g)
    w = torch.randn(cout, cin, k, k, generator=g) / (cin * k * k) ** 0.5
    b = torch.randn(cout, generator=g)
    xi = x.permute(0, 3, 1, 2).double()
    if up == 2:
        xi = F.interpolate(xi, scale_factor=2, mode="nearest")
    ref = F.conv2d(xi, w.double(), b.double(), stride, pad).permute(0, 2, 3, 1)
    cv = _conv_desc(cin, cout, k, stride, pad)
    xd, wd, bd = x.cuda(), w.cuda(), b.cuda()
    out = torch.empty(ref.shape, device="cuda")
    nb = L.cdg_conv2d_workspace_bytes(B, H, H, C.byref(cv), up)
    ws = torch.empty(nb, dtype=torch.uint8, device="cuda")
    s = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    _lib.check(L.cdg_conv2d_forward(0, xd.data_ptr(), B, H, H, wd.data_ptr(), bd.data_ptr(), C.byref(cv), up, out.data_ptr(),
                                    ws.data_ptr(), nb, s))
    torch.cuda.synchronize()
    assert rel(out, ref) < 1e-5


@pytest.mark.parametrize("B,H,cin,cout,k", [(2, 16, 64, 32, 3), (2, 8, 32, 3, 3), (2, 8, 128, 64, 1), (1, 4, 512, 512, 3)])
def test_conv2d_dgrad_matches_torch(B, H, cin, cout, k):
    from cdgvae_b200 import _lib
    L = _lib.lib()
    g = torch.Generator().manual_seed(B + cin + cout + k)
    w = torch.randn(cout, cin, k, k, generator=g) / (cin * k * k) ** 0.5
    go = torch.randn(B, H, H, cout, generator=g)
    x = torch.zeros(B, cin, H, H, dtype=torch.float64, requires_grad=True)
    y = F.conv2d(x, w.double(), None, 1, (k - 1) // 2)
    ref, = torch.autograd.grad(y, x, go.permute(0, 3, 1, 2).double())
    ref = ref.permute(0, 2, 3, 1)
    cv = _conv_desc(cin, cout, k, 1, (k - 1) // 2)
    gd, wd = go.cuda(), w.cuda()
    out = torch.empty(ref.shape, device="cuda")
    nb = L.cdg_conv2d_workspace_bytes(B, H, H, C.byref(cv), 1)
    ws = torch.empty(nb, dtype=torch.uint8, device="cuda")
    s = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    _lib.check(L.cdg_conv2d_dgrad(0, gd.data_ptr(), B, H, H, wd.data_ptr(), C.byref(cv), out.data_ptr(), ws.data_ptr(), nb, s))
    torch.cuda.synchronize()
    assert rel(out, ref) < 1e-5


def _config(scm, batch):
    cfg = dict(node=6, latent_dim=6, scm=scm, flow_num=1, inverse_loop=100, beta=0.1, lr=1e-3, seed=1, batch_size=batch,
               pretrained=False)
    cfg["lambda"] = 5.0
    return cfg


def _full_state(model):
    sd = {k: v.detach().cpu().clone() for k, v in model.state_dict().items()}
    for k, g in enumerate(model.decoder):
        for n, v in g.state_dict().items():
            sd[f"decoder.{k}.{n}"] = v.detach().cpu().clone()
    return sd


def _build(scm, batch):
    from cdgvae_b200.celeba.module.model import CDGVAE
    cfg = _config(scm, batch)
    x, y, n1, n2 = corc.synth_celeba(batch, 1234, 4321)
    masks = torch.split(x[..., 3:], 1, dim=-1)
    torch.manual_seed(cfg["seed"])
    model = CDGVAE(corc.celeba_B(), masks, cfg, "cpu").to("cuda")
    return cfg, model, masks


def adam_param_check(p_new, p_ref, g_ref, lr, tol, what):
    """Updated-parameter parity, as tests/test_pendulum_gpu.py::adam_param_check does it.  Adam's early steps move a weight by
    ~lr * g / (|g| + eps): the SIGN of g decides the update, not its size, so wherever the gradient stands clear of its
    noise floor the updated parameters must agree to `tol` relative even though the gradients themselves only agree to
    GRAD_RTOL.  Elements whose |g| is within 10 % of the tensor's rms (the measured fp32 noise is <= 3 % of rms per element,
    profiles/r02_celeba_grad_noise.txt) may take either sign and are held to Adam's 2 * lr bound instead."""
    p_new, p_ref, g = (t.detach().double().cpu().reshape(-1) for t in (p_new, p_ref, g_ref))
    ill = g.abs() < 0.1 * g.pow(2).mean().sqrt()
    ok = ~ill
    assert float(ok.float().mean()) >= 0.5, (what, "too few well-conditioned elements")
    r = float((p_new[ok] - p_ref[ok]).norm() / (p_ref[ok].norm() + 1e-30))
    assert r < tol, (what, r)
    if ill.any():
        assert float((p_new[ill] - p_ref[ill]).abs().max()) <= 2.2 * lr, what


@pytest.mark.parametrize("name", ["celeba_linear", "celeba_nonlinear", "celeba_b16_linear"])
def test_celeba_step_matches_oracle_and_reference_golden(golden, name):
    from cdgvae_b200.celeba.module.train import train_CDGVAE
    c = golden(name)
    scm, batch = c["config"]["scm"], c["config"]["batch_size"]
    cfg, model, masks = _build(scm, batch)
    # same-seed construction is bit-exact: state_dict keys, trainable set, every tensor
    state = corc.init_state(cfg, cfg["seed"])
    mine = _full_state(model)
    assert set(mine) == set(state)
    for k in state:
        assert torch.equal(mine[k], state[k]), k
    assert sorted(n for n, p in model.named_parameters() if p.requires_grad) == c["trainable"]
    assert not any(k.startswith("decoder") for k in model.state_dict())          # plain list: unregistered (model.py:146)
    opt = torch.optim.Adam(model.parameters(), lr=cfg["lr"])
    spec = corc.CelebaSpec(cfg)
    A = orc.i_b_inv(corc.celeba_B())
    adam = corc.new_adam_state(state)
    for s, e in enumerate(c["steps"], 1):
        x, y, n1, n2 = corc.synth_celeba(batch, 1234 + s - 1, 4321 + s - 1)
        if s > 1:   # one step from identical state: re-synchronise the oracle to the product
            cur = _full_state(model)
            for k in state:
                state[k].copy_(cur[k])
            for n, p in model.named_parameters():
                if p.requires_grad:
                    st = opt.state[p]
                    adam[n]["exp_avg"].copy_(st["exp_avg"].cpu()); adam[n]["exp_avg_sq"].copy_(st["exp_avg_sq"].cpu())
                    adam[n]["step"] = int(st["step"])
        model.noise_fn = None
        q = [n1, n2]
        model.noise_fn = lambda b, d: q.pop(0)
        logs, xhat = train_CDGVAE([(x, y)], model, cfg, opt, "cuda")
        # fp64 oracle on the same inputs from the same state
        st64 = {k: (v.double() if v.dtype.is_floating_point else v.clone()) for k, v in state.items()}
        ad64 = {k: {"step": v["step"], "exp_avg": v["exp_avg"].double(), "exp_avg_sq": v["exp_avg_sq"].double()} for k, v in adam.items()}
        ol, og, oo = corc.train_step(st64, ad64, spec, A.double(), x.double(), y.double(), [m.double() for m in masks], n1.double(), n2.double())
        for k in ("loss", "recon", "KL", "alignment", "active"):
            assert abs(logs[k][0] - ol[k]) <= RTOL * abs(ol[k]) + 1e-7, (s, k, logs[k][0], ol[k])
            if s == 1:
                assert abs(logs[k][0] - e["logs"][k]) <= RTOL * abs(e["logs"][k]) + 1e-7, (s, k, "golden")
        assert rel(xhat, oo["xhat"]) < RTOL
        if s == 1:
            summary_check(xhat, e["xhat"], RTOL, "xhat golden")
        flow_mine, flow_ref = [], []
        for n, p in model.named_parameters():
            if not p.requires_grad:
                assert p.grad is None, n
            elif n.startswith("flows."):
                flow_mine.append(p.grad.reshape(-1).cpu()); flow_ref.append(og[n].reshape(-1))
            else:
                assert rel(p.grad, og[n]) < GRAD_RTOL, (s, n, rel(p.grad, og[n]))
        assert rel(torch.cat(flow_mine), torch.cat(flow_ref)) < GRAD_RTOL, (s, "flows", rel(torch.cat(flow_mine), torch.cat(flow_ref)))
        cur = _full_state(model)
        for k, v in st64.items():
            if k in c["trainable"]:
                # one step from identical state; later steps carry GRAD_RTOL-sized differences of the new gradient into m, v
                adam_param_check(cur[k], v, og[k], cfg["lr"], RTOL if s == 1 else 5 * RTOL, (s, k))
            elif v.dtype.is_floating_point:
                assert rel(cur[k], v) < RTOL, (s, k, rel(cur[k], v))
            else:
                assert torch.equal(cur[k], v), (s, k)                       # num_batches_tracked: +2 encoder, +1 generators
        if s == 1:
            for k, g in e["state"].items():
                if k not in c["trainable"] and not k.endswith("num_batches_tracked"):
                    summary_check(cur[k], g, RTOL, "golden state " + k)
        # carry the oracle's fp32 mirror forward for the next re-synchronisation
        for k in state:
            state[k].copy_(cur[k])


def test_celeba_forward_encode_decode_api():
    cfg, model, masks = _build("linear", 2)
    x, y, n1, n2 = corc.synth_celeba(2, 1234, 4321)
    state = {k: (v.double() if v.dtype.is_floating_point else v.clone()) for k, v in corc.init_state(cfg, 1).items()}
    spec = corc.CelebaSpec(cfg)
    A = orc.i_b_inv(corc.celeba_B()).double()
    ref = corc.forward(dict(state), spec, A, x.double(), [m.double() for m in masks], n1.double(), n2.double())
    q = [n1, n2]
    model.noise_fn = lambda b, d: q.pop(0)
    (mean1, logvar1, eps1, orig, latent, logdet), (mean2, logvar2, eps2), align, sep, xhat = model(x)
    assert logdet == [0] * 6 and len(latent) == 6 and latent[0].shape == (2, 1) and len(align) == 6 and len(sep) == 5
    assert sep[0].shape == (2, 3, 128, 128) and xhat.shape == (2, 128, 128, 3)
    for a, b in ((mean1, ref["mean1"]), (logvar1, ref["logvar1"]), (eps1, ref["epsilon1"]), (orig, ref["orig_latent"]),
                 (torch.cat(latent, 1), torch.cat(ref["latent"], 1)), (mean2, ref["mean2"]), (logvar2, ref["logvar2"]),
                 (eps2, ref["epsilon2"]), (torch.cat(align, 1), torch.cat(ref["align_latent"], 1)), (xhat, ref["xhat"])):
        assert rel(a, b) < RTOL
    for k in range(5):
        assert rel(sep[k], ref["xhat_separated"][k]) < RTOL
    # encode(): the encoder only — generator state must not move; decode(): generators only
    before = _full_state(model)
    q = [n1, n2]
    (m1, _, _, _, lat, _), (_, _, e2) = model.encode(x)
    after = _full_state(model)
    for k in before:
        if k.startswith("decoder."):
            assert torch.equal(before[k], after[k]), k
    assert int(after["encoder.bn1.num_batches_tracked"]) == int(before["encoder.bn1.num_batches_tracked"]) + 1
    sep2, xhat2 = model.decode(lat, e2)
    assert xhat2.shape == (2, 128, 128, 3) and sep2[4].shape == (2, 3, 128, 128)
    mean1b, logvar1b, mean2b, logvar2b = model.get_posterior(x)
    assert rel(mean1b, m1) < 2e-5 and mean2b.shape == (2, 6)      # split-K partials are combined with atomics: run-to-run 1e-6
    model.noise_fn = None
    with pytest.raises(ValueError):
        model(torch.cat([x, x]))                       # masks were taken from a batch of 2 (celeba/main.py:111)
