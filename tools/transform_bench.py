"""Rows/s of the CDG-TVAE data transform kernels on one B200, against the HBM roofline, with the CPU oracle beside it.

    python tools/transform_bench.py [--rows 4194304] [--steps 20] [--warmup 3]

One JSON line per (table shape, direction).  Algorithmic bytes per row:
  forward  8·C (raw fp64) + 8·n_cont (uniforms) + 4·D (transformed fp32)
  inverse  4·D + 8·C (+ 8·n_cont normals + the sigma vector when the draw is on)
Tables are larger than L2 (>= 1.3 GB per launch at the default size); CUDA events around `steps` launches.
"""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rows", type=int, default=1 << 22)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--cpu-rows", type=int, default=20000)
    args = ap.parse_args()
    from oracle import tvae_transform_oracle as orc          # checker + cpu_baseline leg only
    from cdgvae_b200.tabular.modules import data_transformer as DT
    peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {}
    hbm = peaks.get("hbm_gbs", 6550.1)
    for name, n_cont, n_cls in (("loan-shaped (5 continuous)", 5, 0), ("covtype-shaped (7 continuous + Cover_Type)", 7, 7)):
        cols, raw, u, z = orc.synth_table(n_cont, n_cls, 8192, seed=9)
        t = DT.DataTransformer.from_columns(cols)
        C, D = len(cols), t.output_dimensions
        g = torch.Generator(device="cuda").manual_seed(1)
        idx = torch.randint(0, 8192, (args.rows,), device="cuda", generator=g)
        raw_d = torch.from_numpy(raw).cuda()[idx].contiguous()
        ud = torch.rand(n_cont, args.rows, dtype=torch.float64, device="cuda", generator=g)
        zd = torch.randn(n_cont, args.rows, dtype=torch.float64, device="cuda", generator=g)
        sig = torch.full((D,), 0.05, device="cuda")
        out = t.transform(raw_d, uniforms=ud)
        # parity spot check against the oracle on the first rows
        head = orc.transform(cols, raw_d[:2048].cpu().numpy(), ud[:, :2048].cpu().numpy())
        assert np.array_equal(out[:2048].cpu().numpy(), head), "transform differs from the oracle"

        def timed(fn):
            for _ in range(args.warmup):
                fn()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            e0.record()
            for _ in range(args.steps):
                fn()
            e1.record()
            torch.cuda.synchronize()
            return e0.elapsed_time(e1) / args.steps

        # CPU baseline: the oracle (numpy, one core) on a bounded sample
        n = min(args.cpu_rows, args.rows)
        rs, us = raw_d[:n].cpu().numpy(), ud[:, :n].cpu().numpy()
        t0 = time.perf_counter()
        o = orc.transform(cols, rs, us)
        cpu_f = n / (time.perf_counter() - t0)
        t0 = time.perf_counter()
        orc.inverse_transform(cols, o, sig.cpu().numpy(), zd[:, :n].cpu().numpy())
        cpu_i = n / (time.perf_counter() - t0)
        for direction, fn, bytes_row, cpu in (
                ("transform", lambda: t.transform(raw_d, uniforms=ud), 8 * C + 8 * n_cont + 4 * D, cpu_f),
                ("inverse_transform (sigma draw)", lambda: t.inverse_transform(out, sigmas=sig, normals=zd), 4 * D + 8 * C + 8 * n_cont, cpu_i)):
            ms = timed(fn)
            gbs = bytes_row * args.rows / ms / 1e6
            print(json.dumps({"metric": "TVAE data transform rows/sec", "table": name, "direction": direction, "rows": args.rows,
                              "columns": C, "output_dimensions": D, "ms_per_launch": ms, "value": args.rows / ms * 1e3, "unit": "rows/s",
                              "roofline": {"bound": "hbm", "achieved": gbs, "peak": hbm, "unit": "GB/s", "frac": gbs / hbm,
                                           "algorithmic_bytes_per_row": bytes_row},
                              "cpu_baseline": {"value": cpu, "unit": "rows/s", "cores": 1, "kind": "port",
                                               "sample": f"{n} rows through oracle/tvae_transform_oracle.py (numpy)"},
                              "note": "timing includes the output allocation of the public API call"}))


if __name__ == "__main__":
    main()
