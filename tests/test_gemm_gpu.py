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
