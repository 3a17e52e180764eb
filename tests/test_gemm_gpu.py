"""The dense contraction behind every pendulum Linear layer, through cdg_gemm, against a float64
matmul: all operand layouts the step uses (fwd TN, dgrad NN, wgrad with both operands batch-major),
odd sizes, tiny N / K, accumulation."""
import ctypes as C

import pytest
import torch

pytestmark = pytest.mark.gpu

SHAPES = [  # (M, N, K)
    (128, 300, 12288), (128, 300, 300), (128, 8, 300), (128, 300, 1), (128, 300, 2), (16, 300, 192),
    (300, 12288, 128), (3840, 300, 128), (8, 300, 128), (300, 2, 128), (257, 129, 65), (1, 1, 1),
    (1024, 300, 2496), (2496, 300, 1024), (1024, 5952, 300),
]


def run(mode, A, sa, B, sb, M, N, K, accumulate=False, C0=None):
    from cdgvae_b200 import _lib
    out = torch.zeros(M, N, device="cuda") if C0 is None else C0.clone()
    ws = torch.zeros(1 << 20, dtype=torch.uint8, device="cuda")
    s = torch.cuda.current_stream().cuda_stream
    rc = _lib.lib().cdg_gemm(_lib.GEMM_MODES[mode], C.c_void_p(A.data_ptr()), sa[0], sa[1], C.c_void_p(B.data_ptr()),
                             sb[0], sb[1], C.c_void_p(out.data_ptr()), N, M, N, K, int(accumulate),
                             C.c_void_p(ws.data_ptr()), ws.numel(), C.c_void_p(s))
    _lib.check(rc)
    torch.cuda.synchronize()
    return out


@pytest.mark.parametrize("mode", ["simt", "auto", "tc1x", "bf3x"])
@pytest.mark.parametrize("layout", ["tn", "nn", "batch_major"])
@pytest.mark.parametrize("shape", SHAPES)
def test_gemm_layouts(mode, layout, shape):
    M, N, K = shape
    g = torch.Generator().manual_seed(M * 7 + N * 3 + K)
    if layout == "tn":            # A[M,K] k-contiguous, B[N,K] k-contiguous
        A = torch.randn(M, K, generator=g).cuda(); B = torch.randn(N, K, generator=g).cuda()
        sa, sb, ref = (K, 1), (K, 1), A.double() @ B.double().t()
    elif layout == "nn":          # dgrad: B stored [K,N]
        A = torch.randn(M, K, generator=g).cuda(); B = torch.randn(K, N, generator=g).cuda()
        sa, sb, ref = (K, 1), (1, N), A.double() @ B.double()
    else:                          # wgrad: A stored [K,M], B stored [K,N]
        A = torch.randn(K, M, generator=g).cuda(); B = torch.randn(K, N, generator=g).cuda()
        sa, sb, ref = (1, M), (1, N), A.double().t() @ B.double()
    out = run(mode, A, sa, B, sb, M, N, K)
    err = float((out.double() - ref).norm() / ref.norm())
    # 3xTF32: K <= 2048 per accumulation; bf16x3: 16 mantissa bits per operand
    tol = {"tc1x": 2e-3, "simt": 2e-6, "bf3x": 5e-5}.get(mode, 2e-5)
    assert err < tol, (mode, layout, shape, err)
    if layout == "tn" and M * N < 1 << 22:
        C0 = torch.randn(M, N, generator=g).cuda()
        out2 = run(mode, A, sa, B, sb, M, N, K, accumulate=True, C0=C0)
        err2 = float((out2.double() - (ref + C0.double())).norm() / (ref + C0.double()).norm())
        assert err2 < tol


# ---- both operands as bf16 (hi, lo) planes: the CTA-pair kernel of csrc/gemm_ps.cu through cdg_gemm_planes ----------------
def _planes(W, ld16):
    from cdgvae_b200 import _lib
    rows, cols = W.shape
    hi = torch.zeros(rows, ld16, dtype=torch.bfloat16, device="cuda")
    lo = torch.zeros_like(hi)
    s = torch.cuda.current_stream().cuda_stream
    _lib.check(_lib.lib().cdg_split_bf16(C.c_void_p(W.data_ptr()), rows, cols, cols, C.c_void_p(hi.data_ptr()), C.c_void_p(lo.data_ptr()),
                                         ld16, 0, C.c_void_p(s)))
    return hi, lo


PLANE_SHAPES = [(4096, 256, 300), (4096, 304, 300), (4096 + 33, 300, 301), (2048 + 128, 3840, 300), (4096 + 77, 5952, 300), (1024, 2496, 300),
                (4096, 64, 512), (2048, 128, 96), (4096, 304, 1024), (1024, 16, 64), (128, 300, 301), (200, 3840, 301)]


@pytest.mark.parametrize("epi", [0, 1, 2, 3])
@pytest.mark.parametrize("shape", PLANE_SHAPES)
def test_gemm_planes(shape, epi):
    """a = hi + lo carries 16 mantissa bits per operand: 5e-5 against the float64 product of the ORIGINAL fp32 operands (the
    bound the bf16x3 mode has everywhere), 6e-6 against the float64 product of the planes themselves (what the tensor core
    was asked to compute, fp32 accumulation over K <= 2048)."""
    from cdgvae_b200 import _lib
    M, N, K = shape
    g = torch.Generator().manual_seed(M + 3 * N + 7 * K + epi)
    A = torch.randn(M, K, generator=g).cuda(); B = (torch.randn(N, K, generator=g) * 0.1).cuda()
    bias = torch.randn(N, generator=g).cuda(); aux = torch.randn(M, N, generator=g).cuda()
    lda = (K + 15) // 16 * 16
    ah, al = _planes(A, lda); bh, bl = _planes(B, lda)
    out = torch.full((M, N), float("nan"), device="cuda")
    ld_out = (N + 15) // 16 * 16
    oh = torch.zeros(M, ld_out, dtype=torch.bfloat16, device="cuda"); ol = torch.zeros_like(oh)
    s = torch.cuda.current_stream().cuda_stream
    p = lambda t: C.c_void_p(t.data_ptr())
    rc = _lib.lib().cdg_gemm_planes(p(ah), p(al), lda, p(bh), p(bl), lda, p(out), N, M, N, K, epi, p(bias), p(aux), N,
                                    p(oh), p(ol), ld_out, C.c_void_p(s))
    _lib.check(rc)
    torch.cuda.synchronize()

    def post(acc):
        if epi in (1, 2):
            acc = acc + bias.double()
        if epi == 2:
            acc = torch.where(acc > 0, acc, torch.exp(acc) - 1)
        if epi == 3:
            acc = acc * torch.where(aux > 0, torch.ones_like(aux), aux + 1).double()
        return acc
    ref = post(A.double() @ B.double().t())
    Ap = ah[:, :K].double() + al[:, :K].double(); Bp = bh[:, :K].double() + bl[:, :K].double()
    ref_planes = post(Ap @ Bp.t())
    err = float((out.double() - ref).norm() / ref.norm())
    errp = float((out.double() - ref_planes).norm() / ref_planes.norm())
    assert err < 5e-5 and errp < 6e-6, (shape, epi, err, errp)
    # the emitted planes reproduce the fp32 result to 16 mantissa bits
    back = oh[:, :N].double() + ol[:, :N].double()
    assert float((back - out.double()).norm() / out.double().norm()) < 2e-5
    assert torch.equal(oh[:, :N], out.to(torch.bfloat16))                  # hi = round-to-nearest bf16 of the fp32 value


def test_gemm_planes_rejects_what_it_cannot_take():
    from cdgvae_b200 import _lib
    A = torch.randn(64, 64).cuda(); B = torch.randn(32, 64).cuda()
    ah, al = _planes(A, 64); bh, bl = _planes(B, 64)
    out = torch.zeros(64, 32, device="cuda")
    p = lambda t: C.c_void_p(t.data_ptr())
    rc = _lib.lib().cdg_gemm_planes(p(ah), p(al), 64, p(bh), p(bl), 64, p(out), 32, 64, 32, 64, 0, None, None, 0, None, None, 0,
                                    C.c_void_p(torch.cuda.current_stream().cuda_stream))
    assert rc == 4                                                          # CDG_ERR_UNSUPPORTED: M < 128


# ---- long contractions on planes, K-major (input gradients) and MN-major (weight gradients): csrc/gemm_pk.cu ----------------
ACC_SHAPES = [  # (M, N, K, mn_major, with_extra_col)
    (4096, 300, 2496, 0, False), (2048 + 100, 300, 5952, 0, False), (1024, 304, 4096, 0, False), (512, 64, 256, 0, False),
    (2496, 301, 8192, 1, True), (5952, 301, 4096 + 40, 1, True), (300, 300, 16384, 1, False), (1000, 304, 2048, 1, False),
    (300, 301, 8192, 1, True), (128, 16, 64, 1, False),
]


@pytest.mark.parametrize("shape", ACC_SHAPES)
def test_gemm_planes_acc(shape):
    from cdgvae_b200 import _lib
    M, N, K, mn, extra = shape
    g = torch.Generator().manual_seed(M + 3 * N + 7 * K + mn)
    A = torch.randn(M, K, generator=g).cuda(); B = (torch.randn(N, K, generator=g) * 0.1).cuda()
    if extra:
        B[N - 1] = 1.0                                      # the ones column of a weight-gradient's activation operand
    pad8 = lambda n: (n + 7) // 8 * 8
    if mn:                                                  # planes stored [K][M] / [K][N]
        ah, al = _planes(A.t().contiguous(), pad8(M)); bh, bl = _planes(B.t().contiguous(), pad8(N))
        lda, ldb = pad8(M), pad8(N)
        Ap = (ah[:, :M].double() + al[:, :M].double()).t(); Bp = (bh[:, :N].double() + bl[:, :N].double()).t()
    else:
        ah, al = _planes(A, pad8(K)); bh, bl = _planes(B, pad8(K))
        lda = ldb = pad8(K)
        Ap = ah[:, :K].double() + al[:, :K].double(); Bp = bh[:, :K].double() + bl[:, :K].double()
    nc = N - 1 if extra else N
    C0 = torch.randn(M, nc, generator=g).cuda()
    out = C0.clone()
    ex0 = torch.randn(M, generator=g).cuda()
    ex = ex0.clone()
    p = lambda t: C.c_void_p(t.data_ptr())
    rc = _lib.lib().cdg_gemm_planes_acc(p(ah), p(al), lda, p(bh), p(bl), ldb, p(out), nc, M, N, K, mn, p(ex) if extra else None,
                                        C.c_void_p(torch.cuda.current_stream().cuda_stream))
    _lib.check(rc)
    torch.cuda.synchronize()
    full = A.double() @ B.double().t()
    fullp = Ap @ Bp.t()
    ref, refp = C0.double() + full[:, :nc], C0.double() + fullp[:, :nc]
    err = float((out.double() - ref).norm() / ref.norm()); errp = float((out.double() - refp).norm() / refp.norm())
    assert err < 5e-5 and errp < 8e-6, (shape, err, errp)
    if extra:
        r = ex0.double() + full[:, nc]
        assert float((ex.double() - r).norm() / r.norm()) < 5e-5
