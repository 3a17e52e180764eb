"""bf16x3 GEMM with the weight operand pre-split (cdg_split_bf16 + cdg_gemm_bsplit) against the 3xTF32 kernel on the
forward / input-gradient contractions of the pendulum step: TFLOP/s and error against float64.  GPU box only."""
import ctypes as C
import sys

import torch

sys.path.insert(0, ".")
from cdgvae_b200 import _lib  # noqa: E402


def pad8(n):
    return (n + 7) // 8 * 8


def main():
    Bt = int(sys.argv[1]) if len(sys.argv) > 1 else 32768
    L = _lib.lib()
    s = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    P, H = 12288, 300
    g = torch.Generator(device="cuda").manual_seed(0)
    x = torch.randn(Bt, P, device="cuda", generator=g)
    w0 = torch.randn(H, P, device="cuda", generator=g) * 0.01
    h = torch.randn(Bt, H, device="cuda", generator=g)
    w2 = torch.randn(5952, H, device="cuda", generator=g) * 0.05
    gp = torch.randn(Bt, 5952, device="cuda", generator=g)
    ws = torch.zeros(1 << 20, dtype=torch.uint8, device="cuda")
    # (name, A [M,K], W, transpose W?, N, K)   C = A @ Weff^T with Weff[n,k]
    cases = [("enc0_fwd   x @ W0^T", x, w0, 0, H, P), ("dec2_fwd   a2 @ W2^T", h, w2, 0, 5952, H),
             ("dec2_dgrad g @ W2", gp, w2, 1, H, 5952), ("enc1_fwd   h @ W1^T", h, w2[:H].contiguous(), 0, H, H)]
    # weight-gradient contractions: A = the wide streamed operand, batch-major (A(m,k) = A[k, m]); B = the [B,300] operand,
    # split + transposed once.  (name, A [K,M], W=[K,N] source of B, transpose=1, N, K, a_batch_major)
    cases += [("dec2_wgrad g^T @ a2", gp, h, 1, H, Bt), ("enc0_wgrad x^T @ gh", x, h, 1, H, Bt)]
    for name, A, W, tr, N, K in cases:
        bm = name.startswith(("dec2_wgrad", "enc0_wgrad"))
        M = A.shape[1] if bm else A.shape[0]
        ld16 = pad8(K)
        hi = torch.zeros(N, ld16, dtype=torch.bfloat16, device="cuda")
        lo = torch.zeros_like(hi)
        _lib.check(L.cdg_split_bf16(W.data_ptr(), W.shape[0], W.shape[1], W.shape[1], hi.data_ptr(), lo.data_ptr(), ld16, tr, s))
        out = torch.zeros(M, N, device="cuda")
        rows = torch.arange(0, M, max(1, M // 64), device="cuda")[:64]
        Weff = (W.t() if tr else W).double()
        ref = (A[:, rows].t() if bm else A[rows]).double() @ Weff.t()
        sa = (1, A.shape[1]) if bm else (A.shape[1], 1)

        def run_split():
            _lib.check(L.cdg_gemm_bsplit(A.data_ptr(), sa[0], sa[1], hi.data_ptr(), lo.data_ptr(), ld16, out.data_ptr(), N, M, N, K, s))

        def run_mode(mode):
            sb = (1, W.shape[1]) if tr else (W.shape[1], 1)
            _lib.check(L.cdg_gemm(_lib.GEMM_MODES[mode], A.data_ptr(), sa[0], sa[1], W.data_ptr(), sb[0], sb[1], out.data_ptr(), N,
                                  M, N, K, 0, ws.data_ptr(), ws.numel(), s))

        for label, fn in (("tc3x", lambda: run_mode("tc3x")), ("bf3x", lambda: run_mode("bf3x")), ("bf3x+presplit", run_split)):
            out.zero_()
            fn()
            torch.cuda.synchronize()
            err = float((out[rows].double() - ref).norm() / ref.norm())
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            fn(); fn()
            e0.record()
            for _ in range(5):
                fn()
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 5
            print(f"{name:22s} {label:14s} M={M} N={N} K={K}  {ms:7.3f} ms  {2.0 * M * N * K / ms / 1e9:7.1f} TFLOP/s  relerr {err:.2e}", flush=True)


if __name__ == "__main__":
    main()
