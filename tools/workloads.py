"""The secondary workloads of bench.py's `workloads` object (BASELINE.json configs[0], [2], [3], [4]): each runs this
repository's drop-in train loop on synthetic inputs of SURVEY.md §8(d), device-timed with inputs resident in HBM
(`value`), end to end through the same public call with pinned HOST batches (`e2e`), with its own roofline, launch
count and (N = 1, rank 0) the unmodified reference's CPU figure on a bounded sample.  Under torchrun every rank runs
its own shard (weak scaling), gradients exchanged by the train loop itself; times are the max over ranks."""
import os
import sys
from collections import namedtuple

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import synthetic_inputs as syn  # noqa: E402
from tools import reference_arm as ref  # noqa: E402

FP32_PEAK_NOMINAL = 148 * 128 * 2 * 1.965e9 / 1e12     # 74.4 TFLOP/s: 148 SMs x 128 FMA lanes x 2 x 1.965 GHz (not measured)


def _pin(t):
    h = torch.empty(t.shape, dtype=t.dtype, pin_memory=True)
    h.copy_(t)
    return h


class Ctx:
    def __init__(self, dev, world, rank, hbm_gbs, tf_sustained, peak_src):
        self.dev, self.world, self.rank = dev, world, rank
        self.hbm, self.tf, self.src = hbm_gbs, tf_sustained, peak_src

    def barrier(self):
        if self.world > 1:
            import torch.distributed as dist
            dist.barrier()
        torch.cuda.synchronize()

    def max_ms(self, ms):
        t = torch.tensor([ms], device=self.dev)
        if self.world > 1:
            import torch.distributed as dist
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t)

    def timed(self, fn, warm):
        """fn(k) runs k steps; returns ms for the timed call (max over ranks) and the launch count of this rank."""
        from cdgvae_b200 import _lib
        fn(warm)
        self.barrier()
        n0 = _lib.lib().cdg_launch_count()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = fn(None)
        e1.record()
        self.barrier()
        return self.max_ms(e0.elapsed_time(e1)), int(_lib.lib().cdg_launch_count() - n0), out


def _result(c, name, units, unit, steps, warm, ms, launches, e2e_ms, h2d, d2h, roof, cfg, cpu, loss):
    r = {"workload": name, "metric": f"train {unit}", "value": units * c.world * steps / (ms / 1e3), "unit": unit,
         "n_gpus": c.world, "steps": steps, "warmup": warm, "ms_per_step": ms / steps, "gpu_launches": launches,
         "e2e": {"value": units * c.world * steps / (e2e_ms / 1e3), "unit": unit, "ms_per_step": e2e_ms / steps,
                 "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
         "roofline": roof, "config": cfg, "final_loss": loss}
    if cpu is not None:
        r["cpu_baseline"] = cpu
    return r


def tabular(c, dataset="adult", rows=1 << 22, steps=20, warm=6, cpu=True):
    """BASELINE configs[2]: tabular CDG-VAE (tabular/main.py defaults), `rows` rows per GPU per step."""
    from cdgvae_b200.tabular.modules import model as M, train as T
    cfg, mask, ft = ref.tabular_config(dataset)
    cfg["batch_size"] = rows
    torch.manual_seed(1)
    model = M.CDGVAE(syn.tabular_B(dataset), mask, cfg, "cpu").to(c.dev)
    opt = torch.optim.Adam(model.parameters(), lr=cfg["lr"])
    x, y, noise = syn.synth_tabular(dataset, rows, seed=1234 + c.rank)
    xd, yd, nd = x.to(c.dev), y.to(c.dev), noise.to(c.dev)
    ds = namedtuple("DS", ["flatten_topology"])(ft)
    model.noise_fn = lambda n, d: nd
    ms, launches, logs = c.timed(lambda k: T.train_CDGVAE(ds, [(xd, yd)] * (k or steps), model, cfg, opt, c.dev), warm)
    xp, yp, npin = _pin(x), _pin(y), _pin(noise)
    model.noise_fn = lambda n, d: npin
    e2e_ms, _, _ = c.timed(lambda k: T.train_CDGVAE(ds, [(xp, yp)] * (k or steps), model, cfg, opt, c.dev), warm)
    d = cfg["node"]
    bytes_row = 4 * (x.shape[1] + 2 * d)
    ach = rows * steps / (ms / 1e3) * bytes_row / 1e9
    roof = {"bound": "hbm", "achieved": ach, "peak": c.hbm, "unit": "GB/s", "frac": ach / c.hbm, "traffic": None,
            "algorithmic_bytes_per_row": bytes_row, "peak_source": f"{c.src} (MEASURED_PEAKS.json)",
            "note": "nominal bound; the step is fp32-issue bound: ~1,050 instructions per row (244 FFMA, 28 SFU) after the "
                    "constant-bank / register-accumulator rewrite (2,250 before), one row per thread"}
    cpu_r = ref.tabular(dataset, 1 << 16, target_s=4.0, steps=50) if cpu else None
    return _result(c, f"tabular CDG-VAE ({dataset}-shaped table), BASELINE configs[2]", rows, "rows/s", steps, warm, ms, launches,
                   e2e_ms, bytes_row * rows, 4 * (4 + d), roof,
                   {"rows_per_gpu": rows, "dataset": dataset, "parallelism": f"dp{c.world}", "l2_policy":
                    f"inputs {bytes_row * rows / 1e6:.0f} MB per step (> L2 from 2^22 rows); the same device tensors every step, so the "
                    "step's CUDA graph reads them in place"}, cpu_r, logs["loss"][-1])


def tvae(c, kind="loan", rows=1 << 20, steps=10, warm=4, cpu=True):
    """BASELINE configs[3]: CDG-TVAE on a loan- / covtype-shaped transformed table (tabular/main_tvae.py defaults)."""
    from cdgvae_b200.tabular.modules import model as M, train as T
    cfg, oil, mask, Bm = ref.tvae_config(kind)
    cfg["batch_size"] = rows
    torch.manual_seed(1)
    model = M.TVAE(Bm, mask, cfg, "cpu").to(c.dev)
    opt = torch.optim.Adam(model.parameters(), lr=cfg["lr"], weight_decay=cfg["weight_decay"])
    x, y, noise = syn.synth_tvae(kind, rows, seed=1234 + c.rank)
    xd, yd, nd = x.to(c.dev), y.to(c.dev), noise.to(c.dev)
    Span = namedtuple("SpanInfo", ["dim", "activation_fn"])
    o = [[Span(*s) for s in col] for col in oil]
    model.noise_fn = lambda n, d: nd
    ms, launches, logs = c.timed(lambda k: T.train_TVAE(o, None, [(xd, yd)] * (k or steps), model, cfg, opt, c.dev), warm)
    xp, yp, npin = _pin(x), _pin(y), _pin(noise)
    model.noise_fn = lambda n, d: npin
    e2e_ms, _, _ = c.timed(lambda k: T.train_TVAE(o, None, [(xp, yp)] * (k or steps), model, cfg, opt, c.dev), warm)
    d, D = cfg["node"], cfg["input_dim"]
    bytes_row = 4 * (D + 2 * d)
    macs = 32 * D + 512 + 256 + 32 * d + sum(8 * k + 64 + 128 + 16 * m for k, m in zip(cfg["factor"], mask))
    flops_row = 6.0 * macs
    rows_s = rows * steps / (ms / 1e3)
    ach = rows_s * bytes_row / 1e9
    roof = {"bound": "hbm", "achieved": ach, "peak": c.hbm, "unit": "GB/s", "frac": ach / c.hbm, "traffic": None,
            "algorithmic_bytes_per_row": bytes_row, "algorithmic_flops_per_row": flops_row,
            "fp32_tflops": rows_s * flops_row / 1e12, "frac_of_fp32_peak_nominal": rows_s * flops_row / 1e12 / FP32_PEAK_NOMINAL,
            "peak_source": f"{c.src} (MEASURED_PEAKS.json); fp32 peak nominal {FP32_PEAK_NOMINAL:.1f} TFLOP/s",
            "note": "HBM bound stated as BASELINE.md does; in practice fp32 compute bound (SURVEY §8d row 4)"}
    cpu_r = ref.tvae(kind, 1 << 15, target_s=4.0, steps=50) if cpu else None
    return _result(c, f"CDG-TVAE ({kind}-shaped table, D={D}), BASELINE configs[3]", rows, "rows/s", steps, warm, ms, launches, e2e_ms,
                   bytes_row * rows, 4 * (4 + d), roof, {"rows_per_gpu": rows, "shape": kind, "D": D, "parallelism": f"dp{c.world}"},
                   cpu_r, logs["loss"][-1])


def celeba(c, batch=16, steps=10, warm=6, cpu=True):
    """BASELINE configs[4]: CelebA-shaped CDG-VAE, the reference's batch 16 per GPU (celeba/main.py:70)."""
    from cdgvae_b200.celeba.module.model import CDGVAE
    from cdgvae_b200.celeba.module.train import train_CDGVAE
    cfg = ref.celeba_config(batch)
    x, y, n1, n2 = syn.synth_celeba(batch, seed=1234 + c.rank)
    xd, yd = x.to(c.dev), y.to(c.dev)
    torch.manual_seed(1)
    model = CDGVAE(syn.celeba_B(), torch.split(xd[..., 3:], 1, dim=-1), cfg, c.dev)
    opt = torch.optim.Adam(model.parameters(), lr=cfg["lr"])
    nd = n1.to(c.dev)
    model.noise_fn = lambda b, d: nd
    ms, launches, (logs, _) = c.timed(lambda k: train_CDGVAE([(xd, yd)] * (k or steps), model, cfg, opt, c.dev), warm)
    xp, yp, npin = _pin(x), _pin(y), _pin(n1)
    model.noise_fn = lambda b, d: npin
    e2e_ms, _, _ = c.timed(lambda k: train_CDGVAE([(xp, yp)] * (k or steps), model, cfg, opt, c.dev), warm)
    flops = 47.6e9 * batch
    ach = flops * steps / (ms / 1e3) / 1e12
    roof = {"bound": "tensor", "achieved": ach, "peak": c.tf, "unit": "TFLOP/s", "frac": ach / c.tf, "traffic": None,
            "algorithmic_flops_per_sample": 47.6e9, "frac_of_fp32_faithful_ceiling": ach / (c.tf / 6.0),
            "peak_source": f"bf16 dense sustained, {c.src} (MEASURED_PEAKS.json)",
            "note": "3xTF32 arithmetic (6 bf16-equivalent MMAs per product): frac <= 1/6 by construction"}
    cpu_r = ref.celeba(batch, warmup=1, steps=1) if cpu else None
    return _result(c, "CelebA-shaped CDG-VAE (128x128, ResNet-18 encoder, 5 SAGAN generators), BASELINE configs[4]", batch, "samples/s",
                   steps, warm, ms, launches, e2e_ms, x.numel() * 4 + y.numel() * 4 + 2 * n1.numel() * 4, 4 * 5, roof,
                   {"batch_per_gpu": batch, "parallelism": f"dp{c.world}", "batchnorm": "per-shard statistics (SURVEY §8e)"},
                   cpu_r, logs["loss"][-1])


def pendulum_b128(c, batch=128, steps=60, warm=10, cpu=True):
    """BASELINE configs[0] on the GPU: train_CDGVAE at the reference's own batch 128 (main.py:96), linear SCM."""
    from cdgvae_b200.modules.model import CDGVAE
    from cdgvae_b200.modules.train import train_CDGVAE
    cfg = ref.pendulum_config(batch, 0, "linear")
    torch.manual_seed(1)
    model = CDGVAE(syn.pendulum_B(4), syn.pendulum_masks(64), cfg, "cpu").to(c.dev)
    opt = torch.optim.Adam(model.parameters(), lr=cfg["lr"])
    x, y, noise = syn.synth_pendulum(batch, 64, 4, seed=1234 + c.rank)
    xd, yd, nd = x.to(c.dev), y.to(c.dev), noise.to(c.dev)
    model.noise_fn = lambda n, d: nd
    ms, launches, (logs, _) = c.timed(lambda k: train_CDGVAE([(xd, yd)] * (k or steps), model, cfg, opt, c.dev), warm)
    xp, yp, npin = _pin(x), _pin(y), _pin(noise)
    model.noise_fn = lambda n, d: npin
    e2e_ms, _, _ = c.timed(lambda k: train_CDGVAE([(xp, yp)] * (k or steps), model, cfg, opt, c.dev), warm)
    live = 7_751_104
    bytes_step = 32 * live + 49_188 * batch
    ach = bytes_step * steps / (ms / 1e3) / 1e9
    roof = {"bound": "hbm", "achieved": ach, "peak": c.hbm, "unit": "GB/s", "frac": ach / c.hbm, "traffic": None,
            "algorithmic_bytes_per_step": bytes_step, "peak_source": f"{c.src} (MEASURED_PEAKS.json)",
            "note": "small-batch regime: weights + optimizer traffic dominate (32 B x 7,751,104 live parameters per step)"}
    cpu_r = ref.pendulum_b128(batch, target_s=4.0, steps=60) if cpu else None
    return _result(c, "pendulum CDG-VAE at the reference's batch 128 (main.py --model CDGVAE), BASELINE configs[0]", batch, "samples/s",
                   steps, warm, ms, launches, e2e_ms, x.numel() * 4 + y.numel() * 4 + noise.numel() * 4, 4 * 8, roof,
                   {"batch_per_gpu": batch, "scm": "linear", "parallelism": f"dp{c.world}"}, cpu_r, logs["loss"][-1])
