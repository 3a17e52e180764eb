"""Parity at sizes the oracle cannot reach in seconds, through size-independent properties of the step
(BASELINE full-size configs: pendulum 64x64x3, batches in the thousands; tabular 2^18 rows):

  * mean-of-shards: every loss is a batch mean of per-sample terms, so loss/gradients of a batch equal the
    average over its two halves (this is also what makes the path data-parallel);
  * the tensor-core path (3xTF32, fused reconstruction epilogue, split-K) against the fp32 SIMT path on the
    same large batch;
  * a batch made of one sample repeated gives that sample's loss and gradient;
  * dead decoder columns stay exactly zero at full size.
"""
import pytest
import torch

from oracle import cdgvae_oracle as orc

pytestmark = pytest.mark.gpu


def rel(a, b):
    a, b = a.double().reshape(-1), b.double().reshape(-1)
    return float((a - b).norm() / (b.norm() + 1e-30))


def pend_model(mode, scm="nonlinear"):
    from cdgvae_b200.modules.model import CDGVAE
    cfg = dict(node=4, scm=scm, flow_num=1, inverse_loop=100, factor=[1, 1, 2], image_size=64, lr=1e-3, beta=0.1,
               gemm_mode=mode)
    cfg["lambda"] = 5.0
    torch.manual_seed(1)
    return CDGVAE(orc.pendulum_B(4), orc.pendulum_masks(64), cfg, "cpu").to("cuda"), cfg


def fwd_bwd(model, x, y, noise, x_l=None, y_l=None):
    row = torch.zeros(8, device="cuda")
    model.forward_backward(x, y, noise, row, x_l=x_l, y_l=y_l)
    torch.cuda.synchronize()
    return row.clone(), model._grads.clone()


def synth(B, seed):
    g = torch.Generator(device="cuda").manual_seed(seed)
    x = torch.rand(B, 64, 64, 3, device="cuda", generator=g) * 2 - 1
    x.masked_fill_(torch.rand(B, 64, 64, 3, device="cuda", generator=g) < 0.9, 1.0)
    return x, torch.rand(B, 5, device="cuda", generator=g), torch.randn(B, 4, device="cuda", generator=g)


@pytest.mark.parametrize("B", [4096])
def test_pendulum_mean_of_shards_and_tc_vs_simt(B):
    x, y, noise = synth(B, 7)
    tc, _ = pend_model("auto")
    simt, _ = pend_model("simt")
    row_tc, g_tc = fwd_bwd(tc, x, y, noise)
    row_si, g_si = fwd_bwd(simt, x, y, noise)
    # tensor-core path vs fp32 CUDA-core path, full-size batch
    assert rel(row_tc, row_si) < 1e-4
    for n, p in tc.named_parameters():
        o, k = tc._offsets[n], p.numel()
        assert rel(g_tc[o:o + k], g_si[o:o + k]) < 1e-4, n
    # mean of shards
    h = B // 2
    r0, g0 = fwd_bwd(tc, x[:h], y[:h], noise[:h])
    r1, g1 = fwd_bwd(tc, x[h:], y[h:], noise[h:])
    assert rel((r0 + r1) / 2, row_tc) < 1e-4
    assert rel((g0 + g1) / 2, g_tc) < 1e-4
    # dead decoder output rows: exactly zero gradient
    for k, (lo, hi) in enumerate(tc._ranges):
        o = tc._offsets[f"decoder.{k}.4.weight"]
        w = g_tc[o:o + 12288 * 300].view(12288, 300)
        assert float(w[:lo].abs().max() if lo else 0.0) == 0.0 and float(w[hi:].abs().max() if hi < 12288 else 0.0) == 0.0


def test_pendulum_semi_repeated_sample_is_size_independent():
    """A batch that repeats one (x, noise) row B times has the loss / gradient of that single row."""
    x, y, noise = synth(2, 11)
    xl, yl, _ = synth(2, 12)
    model, _ = pend_model("auto")
    r1, g1 = fwd_bwd(model, x[:1].repeat(256, 1, 1, 1), None, noise[:1].repeat(256, 1), x_l=xl[:1].repeat(64, 1, 1, 1),
                     y_l=yl[:1].repeat(64, 1))
    r2, g2 = fwd_bwd(model, x[:1].repeat(2048, 1, 1, 1), None, noise[:1].repeat(2048, 1), x_l=xl[:1].repeat(512, 1, 1, 1),
                     y_l=yl[:1].repeat(512, 1))
    assert rel(r2, r1) < 1e-4
    assert rel(g2, g1) < 1e-4


@pytest.mark.parametrize("kind", ["adult", "covtype", "tvae_loan"])
def test_tabular_mean_of_shards_large(kind):
    from cdgvae_b200.tabular.modules import model as M
    n = 1 << 18
    if kind.startswith("tvae"):
        oil, mask, d, Bm, D = orc.tvae_shape("loan")
        cfg = dict(dataset="loan", scm="linear", flow_num=1, inverse_loop=100, node=d, factor=[1] * d, input_dim=D,
                   sigma_range=[0.01, 0.1])
        cfg["lambda"] = 5.0
        torch.manual_seed(1)
        model = M.TVAE(Bm, mask, cfg, "cpu").to("cuda")
        x, y, noise = orc.synth_tvae("loan", n)
        aux = dict(output_info_list=oil)
    else:
        d = 6 if kind == "covtype" else 3
        cfg = dict(dataset=kind, scm="linear", flow_num=1, inverse_loop=100, beta=0.01, node=d, factor=[1] * d,
                   input_dim=8 if kind == "covtype" else 5)
        cfg["lambda"] = 10.0
        mask = [1, 1, 2, 1, 1, 8] if kind == "covtype" else [1, 1, 3]
        torch.manual_seed(1)
        model = M.CDGVAE(orc.tabular_B(kind), mask, cfg, "cpu").to("cuda")
        x, y, noise = orc.synth_tabular(kind, n)
        aux = dict(flatten_topology=None if kind == "covtype" else [2, 3, 0, 1, 4])
    x, y, noise = x.cuda(), y.cuda(), noise.cuda()

    def run(sl):
        row = torch.zeros(4 + d, device="cuda")
        model.forward_backward(x[sl], y[sl], noise[sl], row, **aux)
        torch.cuda.synchronize()
        return row.clone(), model._grads.clone()
    rf, gf = run(slice(0, n))
    r0, g0 = run(slice(0, n // 2))
    r1, g1 = run(slice(n // 2, n))
    assert rel((r0 + r1) / 2, rf) < 1e-4
    assert rel((g0 + g1) / 2, gf) < 2e-4
    # against the oracle on a prefix the CPU finishes quickly: same per-row arithmetic at any batch size
    m = 4096
    spec = orc.tvae_spec(cfg, mask, aux["output_info_list"]) if kind.startswith("tvae") else orc.tabular_spec(cfg, mask, aux["flatten_topology"])
    params = {k: v.detach().cpu().clone() for k, v in model.state_dict().items()}
    leaves = {k: v.requires_grad_(True) for k, v in params.items()}
    A = orc.i_b_inv(model.B.cpu())
    loss, logs, _ = orc.step_losses(leaves, spec, A, x[:m].cpu(), y[:m].cpu(), noise[:m].cpu())
    rp, _ = run(slice(0, m))
    assert abs(float(rp[0]) - float(loss.detach())) <= 1e-4 * abs(float(loss.detach()))
