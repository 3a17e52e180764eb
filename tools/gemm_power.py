"""SM clock / board power while ONE pre-split bf16x3 GEMM (encoder first layer forward) runs back to back for a few seconds:
tells a power-capped tensor pipe from a stalled one.  GPU box only.   python tools/gemm_power.py [seconds]"""
import ctypes as C
import sys
import threading
import time

import torch

sys.path.insert(0, ".")
from cdgvae_b200 import _lib  # noqa: E402


def main():
    secs = float(sys.argv[1]) if len(sys.argv) > 1 else 3.0
    import pynvml
    pynvml.nvmlInit()
    h = pynvml.nvmlDeviceGetHandleByIndex(0)
    L = _lib.lib()
    s = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    M, N, K = 32768, 300, 12288
    g = torch.Generator(device="cuda").manual_seed(0)
    x = torch.randn(M, K, device="cuda", generator=g)
    w = torch.randn(N, K, device="cuda", generator=g) * 0.01
    hi = torch.zeros(N, K, dtype=torch.bfloat16, device="cuda")
    lo = torch.zeros_like(hi)
    _lib.check(L.cdg_split_bf16(w.data_ptr(), N, K, K, hi.data_ptr(), lo.data_ptr(), K, 0, s))
    out = torch.zeros(M, N, device="cuda")
    samples, stop = [], [False]

    def sample():
        while not stop[0]:
            samples.append((pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM), pynvml.nvmlDeviceGetPowerUsage(h) / 1e3))
            time.sleep(0.05)

    def run():
        _lib.check(L.cdg_gemm_bsplit(x.data_ptr(), K, 1, hi.data_ptr(), lo.data_ptr(), K, out.data_ptr(), N, M, N, K, s))

    for _ in range(3):
        run()
    torch.cuda.synchronize()
    t = threading.Thread(target=sample)
    t.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0, n = time.time(), 0
    e0.record()
    while time.time() - t0 < secs:
        for _ in range(50):
            run()
        n += 50
        torch.cuda.synchronize()
    e1.record()
    torch.cuda.synchronize()
    stop[0] = True
    t.join()
    ms = e0.elapsed_time(e1) / n
    tail = samples[len(samples) // 2:]
    mhz = sorted(c for c, _ in tail)[len(tail) // 2]
    pw = sorted(p for _, p in tail)[len(tail) // 2]
    print(f"enc0_fwd M={M}: {ms:.3f} ms  {2.0 * M * N * K / ms / 1e9:.1f} TFLOP/s  sm {mhz} MHz  power {pw:.0f} W (medians of the second half, {len(tail)} samples)")


if __name__ == "__main__":
    main()
