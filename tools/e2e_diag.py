"""Where the end-to-end step time goes: host->device copy alone, copy + byte conversion, and the full e2e loop.

    python tools/e2e_diag.py [--batch 131072] [--steps 5]
"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=1 << 17)
    ap.add_argument("--steps", type=int, default=5)
    args = ap.parse_args()
    import bench
    from oracle import cdgvae_oracle as orc
    from cdgvae_b200.data import DevicePrefetcher
    from cdgvae_b200.modules.model import CDGVAE
    from cdgvae_b200.modules.train import train_CDGVAE_semi_loaders
    dev = torch.device("cuda", 0)
    B, BL = args.batch, args.batch // 4
    cfg = bench.make_config(B, BL)
    torch.manual_seed(1)
    model = CDGVAE(orc.pendulum_B(4), orc.pendulum_masks(64), cfg, "cpu").to(dev)
    opt = torch.optim.Adam(model.parameters(), lr=cfg["lr"])
    xd, xld, yld, nd, xu8, xlu8 = bench.synth_device(B, BL, 1234, dev)
    model.noise_fn = lambda n, d: nd

    def pinned(t):
        h = torch.empty(t.shape, dtype=t.dtype, pin_memory=True)
        h.copy_(t)
        return h
    xp, xlp, ylp = pinned(xu8), pinned(xlu8), pinned(yld)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

    def timed(fn):
        fn(2)
        torch.cuda.synchronize()
        e0.record()
        fn(args.steps)
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / args.steps

    def drain(pixels):
        def f(k):
            for _ in zip(DevicePrefetcher([(xlp, ylp)] * k, dev, pixels=pixels), DevicePrefetcher([xp] * k, dev, pixels=pixels)):
                pass
        return f
    out = {"batch": B}
    out["copy_only_ms"] = timed(drain(False))
    out["copy_convert_ms"] = timed(drain(True))
    out["compute_only_ms"] = timed(lambda k: train_CDGVAE_semi_loaders([(xld, yld)] * k, [xd] * k, model, cfg, opt, dev))
    for mode in ("copy", "own", "compute"):
        out["e2e_cold_ms_convert_on_" + mode] = timed(lambda k: train_CDGVAE_semi_loaders(
            DevicePrefetcher([(xlp, ylp)] * k, dev, pixels=True, convert_on=mode),
            DevicePrefetcher([xp] * k, dev, pixels=True, convert_on=mode), model, cfg, opt, dev))
    # steady state: 12 steps in one loop, clock from step 4
    class Marked:
        exact_len = True
        def __init__(self, loader, at): self.loader, self.at = loader, at
        def __len__(self): return len(self.loader)
        def __iter__(self):
            for i, b in enumerate(self.loader):
                if i == self.at: e0.record()
                yield b
    for mode in ("copy", "own", "compute"):
        k = 12
        torch.cuda.synchronize()
        train_CDGVAE_semi_loaders(DevicePrefetcher([(xlp, ylp)] * k, dev, pixels=True, convert_on=mode),
                                  Marked(DevicePrefetcher([xp] * k, dev, pixels=True, convert_on=mode), 4), model, cfg, opt, dev)
        e1.record()
        torch.cuda.synchronize()
        out["e2e_warm_ms_convert_on_" + mode] = e0.elapsed_time(e1) / (k - 4)
    # the same with the images already fp32 on the host and with U-only uint8 (no labeled loader H2D)
    # compute while an unrelated H2D stream is running: does the copy slow the step, or the step the copy?
    side = torch.cuda.Stream(dev)
    sink = torch.empty_like(xu8)
    def with_bg(k):
        with torch.cuda.stream(side):
            for _ in range(k):
                sink.copy_(xp, non_blocking=True)
        train_CDGVAE_semi_loaders([(xld, yld)] * k, [xd] * k, model, cfg, opt, dev)
        torch.cuda.current_stream().wait_stream(side)
    out["compute_with_background_copy_ms"] = timed(with_bg)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
