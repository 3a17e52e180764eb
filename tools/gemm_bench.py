"""Time cdg_gemm on the five large contractions of the pendulum step in every mode and
report TFLOP/s and the error against float64 on a row sample.  GPU box only."""
import ctypes as C
import sys

import torch

sys.path.insert(0, ".")
from cdgvae_b200 import _lib  # noqa: E402


def run(mode, A, sa, B, sb, M, N, K, out, ws):
    s = torch.cuda.current_stream().cuda_stream
    rc = _lib.lib().cdg_gemm(_lib.GEMM_MODES[mode], C.c_void_p(A.data_ptr()), sa[0], sa[1], C.c_void_p(B.data_ptr()),
                             sb[0], sb[1], C.c_void_p(out.data_ptr()), N, M, N, K, 0, C.c_void_p(ws.data_ptr()),
                             ws.numel(), C.c_void_p(s))
    _lib.check(rc)


def main():
    Bt = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
    modes = sys.argv[2].split(",") if len(sys.argv) > 2 else ["simt", "tc3x", "tc1x"]
    P, H = 12288, 300
    ws = torch.zeros(1 << 20, dtype=torch.uint8, device="cuda")
    g = torch.Generator(device="cuda").manual_seed(0)
    x = torch.randn(Bt, P, device="cuda", generator=g)
    w0 = torch.randn(H, P, device="cuda", generator=g) * 0.01
    h = torch.randn(Bt, H, device="cuda", generator=g)
    w2 = torch.randn(5952, H, device="cuda", generator=g) * 0.05
    gp = torch.randn(Bt, 5952, device="cuda", generator=g)
    cases = [
        ("enc0_fwd  x[B,P] W0[H,P]^T", x, (P, 1), w0, (P, 1), Bt, H, P),
        ("dec2_fwd  a2[B,H] W2[N,H]^T", h, (H, 1), w2, (H, 1), Bt, 5952, H),
        ("dec2_dgrad g[B,N] W2[N,H]", gp, (5952, 1), w2, (1, H), Bt, H, 5952),
        ("dec2_wgrad g[B,N]^T a2[B,H]", gp, (1, 5952), h, (1, H), 5952, H, Bt),
        ("enc0_wgrad gh[B,H]^T x[B,P]", h, (1, H), x, (1, P), H, P, Bt),
        ("enc1_fwd  h[B,H] W1[H,H]^T", h, (H, 1), w2[:H].contiguous(), (H, 1), Bt, H, H),
    ]
    for name, A, sa, B, sb, M, N, K in cases:
        out = torch.zeros(M, N, device="cuda")
        # float64 reference on a sample of rows
        rows = torch.arange(0, M, max(1, M // 64), device="cuda")[:64]
        Ad = (A[rows] if sa[1] == 1 else A[:, rows].t()).double()
        Bd = (B if sb[1] == 1 else B.t()).double()
        ref = Ad @ Bd.t()
        for mode in modes:
            run(mode, A, sa, B, sb, M, N, K, out, ws)
            torch.cuda.synchronize()
            err = float((out[rows].double() - ref).norm() / ref.norm())
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            for _ in range(2):
                run(mode, A, sa, B, sb, M, N, K, out, ws)
            e0.record()
            reps = 5
            for _ in range(reps):
                run(mode, A, sa, B, sb, M, N, K, out, ws)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / reps
            print(f"{name:32s} {mode:5s} M={M:6d} N={N:6d} K={K:6d}  {ms:8.3f} ms  {2.0 * M * N * K / ms / 1e9:8.1f} TFLOP/s  relerr {err:.2e}",
                  flush=True)


if __name__ == "__main__":
    main()
