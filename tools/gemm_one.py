"""Run one large contraction through cdg_gemm a few times (for ncu captures).  usage: gemm_one.py <case> <mode> [B]"""
import ctypes as C
import sys

import torch

sys.path.insert(0, ".")
from cdgvae_b200 import _lib  # noqa: E402

case, mode = sys.argv[1], sys.argv[2]
Bt = int(sys.argv[3]) if len(sys.argv) > 3 else 32768
P, H = 12288, 300
g = torch.Generator(device="cuda").manual_seed(0)
ws = torch.zeros(1 << 20, dtype=torch.uint8, device="cuda")
if case == "enc0_fwd":
    A = torch.randn(Bt, P, device="cuda", generator=g); Bm = torch.randn(H, P, device="cuda", generator=g)
    sa, sb, M, N, K = (P, 1), (P, 1), Bt, H, P
elif case == "dec2_dgrad":
    A = torch.randn(Bt, 5952, device="cuda", generator=g); Bm = torch.randn(5952, H, device="cuda", generator=g)
    sa, sb, M, N, K = (5952, 1), (1, H), Bt, H, 5952
elif case == "enc0_wgrad":
    A = torch.randn(Bt, H, device="cuda", generator=g); Bm = torch.randn(Bt, P, device="cuda", generator=g)
    sa, sb, M, N, K = (1, H), (1, P), H, P, Bt
else:
    A = torch.randn(Bt, H, device="cuda", generator=g); Bm = torch.randn(5952, H, device="cuda", generator=g)
    sa, sb, M, N, K = (H, 1), (H, 1), Bt, 5952, H
out = torch.zeros(M, N, device="cuda")
s = torch.cuda.current_stream().cuda_stream
for _ in range(3):
    _lib.check(_lib.lib().cdg_gemm(_lib.GEMM_MODES[mode], C.c_void_p(A.data_ptr()), sa[0], sa[1], C.c_void_p(Bm.data_ptr()),
                                   sb[0], sb[1], C.c_void_p(out.data_ptr()), N, M, N, K, 0, C.c_void_p(ws.data_ptr()),
                                   ws.numel(), C.c_void_p(s)))
torch.cuda.synchronize()
print("done", float(out.abs().mean()))
