import torch, time
n = 1 << 30   # 4 GiB of fp32
h = torch.empty(n, dtype=torch.float32, pin_memory=True); h.fill_(1.0)
d = torch.empty(n, dtype=torch.float32, device="cuda")
for chunks, streams in ((1, 1), (4, 1), (4, 2), (8, 4)):
    ss = [torch.cuda.Stream() for _ in range(streams)]
    torch.cuda.synchronize(); t0 = time.perf_counter()
    step = n // chunks
    for i in range(chunks):
        with torch.cuda.stream(ss[i % streams]):
            d[i * step:(i + 1) * step].copy_(h[i * step:(i + 1) * step], non_blocking=True)
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    print(f"H2D {chunks} chunks on {streams} streams: {4 * n / dt / 1e9:.1f} GB/s")
