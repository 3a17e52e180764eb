#!/usr/bin/env python
"""bench.py — CDG-VAE training throughput on B200 (BASELINE.json metric).

Workload at N=1 (BASELINE.json configs[1]): pendulum CDG-VAE, semi-supervised step
(main_semi.py defaults: scm nonlinear, flow_num 1, beta 0.1, lambda 5, lr 1e-3, labeled batch =
unlabeled/4), synthetic pendulum-shaped images, large per-GPU batch.  A "step" is one pass of
train_CDGVAE_semi's loop body over one batch: zero_grad, forward, losses, backward, Adam.

    python bench.py --gpus 1 --steps 5 --warmup 3
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...
    python bench.py --impl reference          # the CPU baseline arm (oracle port on host cores)
"""
import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

MACS_U = 19_522_800      # fwd 7,736,400 + wgrad 7,736,400 + dgrad 4,050,000 per unlabeled sample
MACS_L = 3_778_800 * 2 + 92_400   # labeled sample: encoder fwd + wgrad + dgrad (L1, L2)


# dram__bytes_read.sum + dram__bytes_write.sum of ONE launch of the dominant kernel, from the committed
# `ncu --set full` capture (profiles/r01_ncu_enc0_fwd_bf3x_presplit.txt): encoder first Linear forward,
# x[32768,12288] @ W0[300,12288]^T.  Algorithmic bytes of that launch: x once + W0 once (as bf16 hi + lo) + h1 written once.
NCU_TRAFFIC = {"bytes": 1.899331e9 + 47.327488e6,   # profiles/r01_ncu_gemm_cta2_final.txt, launch 0
               "launch": "gemm_tc_kernel<304,32,2,0,0,0,1> (bf16x3, pre-split weights, CTA pairs) enc0 forward, M=32768 N=300 K=12288",
               "algorithmic": 32768 * 12288 * 4 + 300 * 12288 * 4 + 32768 * 300 * 4}


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d["hbm_gbs"], d["bf16_tflops"], d.get("bf16_tflops_sustained", d["bf16_tflops"]), "measured"
    return 6650.0, 1590.0, 1400.0, "fallback"


def make_config(batch, batch_l):
    cfg = dict(node=4, scm="nonlinear", flow_num=1, inverse_loop=100, factor=[1, 1, 2], image_size=64,
               batch_size=batch, batch_sizeL=batch_l, lr=1e-3, beta=0.1, seed=1, labeled_ratio=0.1)
    cfg["lambda"] = 5.0
    return cfg


class ClockSampler(threading.Thread):
    """nvidia-smi style clock / throttle-reason sampling during the timed region (NVML)."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.stop_flag, self.max_mhz = index, [], set(), False, None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def run(self):
        if self.nv is None:
            return
        nv = self.nv
        names = {nv.nvmlClocksThrottleReasonSwPowerCap: "sw_power_cap",
                 nv.nvmlClocksThrottleReasonHwSlowdown: "hw_slowdown",
                 nv.nvmlClocksThrottleReasonSwThermalSlowdown: "sw_thermal_slowdown",
                 nv.nvmlClocksThrottleReasonHwThermalSlowdown: "hw_thermal_slowdown",
                 nv.nvmlClocksThrottleReasonHwPowerBrakeSlowdown: "hw_power_brake"}
        while not self.stop_flag:
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, nm in names.items():
                    if r & bit:
                        self.reasons.add(nm)
            except Exception:
                pass
            time.sleep(0.05)

    def result(self):
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2] if s else None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(s)}


def synth(batch, batch_l, seed):
    """Synthetic pendulum-shaped batch: 90 % white pixels, rest U(-1,1); y ~ U(0,1)."""
    g = torch.Generator().manual_seed(seed)
    x = torch.empty(batch, 64, 64, 3)
    x.uniform_(-1, 1, generator=g)
    x[torch.rand(batch, 64, 64, 3, generator=g) < 0.9] = 1.0
    xl = x[:batch_l].clone().roll(1, 0)
    yl = torch.rand(batch_l, 5, generator=g)
    noise = torch.randn(batch, 4, generator=g)
    return x, xl, yl, noise


def synth_device(batch, batch_l, seed, dev):
    """Same distribution as synth(), generated on the device (a 131072-image batch is 6.4 GB: building it on the
    host would cost tens of seconds and a second host copy per rank)."""
    from cdgvae_b200 import _lib
    g = torch.Generator(device=dev).manual_seed(seed)
    # image bytes as the pendulum PNGs hold them (modules/datasets.py:24-27): 90 % white (255), the rest uniform
    xu8 = torch.randint(0, 256, (batch, 64, 64, 3), dtype=torch.uint8, device=dev, generator=g)
    xu8.masked_fill_(torch.rand(batch, 64, 64, 3, device=dev, generator=g) < 0.9, 255)
    x = torch.empty(batch, 64, 64, 3, device=dev)
    _lib.check(_lib.lib().cdg_pixels_to_float(xu8.data_ptr(), xu8.numel(), x.data_ptr(),          # datasets.py:28
                                              torch.cuda.current_stream(dev).cuda_stream))
    xlu8 = xu8[:batch_l].roll(1, 0).contiguous()
    xl = x[:batch_l].roll(1, 0).contiguous()
    yl = torch.rand(batch_l, 5, device=dev, generator=g)
    noise = torch.randn(batch, 4, device=dev, generator=g)
    return x, xl, yl, noise, xu8, xlu8


def cpu_baseline(threads, target_s=12.0, batch=1024, batch_l=256):
    """The oracle port of the reference step, timed on the host cores (a reported baseline)."""
    from oracle import cdgvae_oracle as orc
    torch.set_num_threads(threads)
    cfg = make_config(batch, batch_l)
    mask = orc.pendulum_masks(64)
    spec = orc.pendulum_spec(cfg, mask)
    A = orc.i_b_inv(orc.pendulum_B(4))
    params = orc.init_params(spec, 1)
    adam = orc.new_adam_state(params)
    x, xl, yl, noise = synth(batch, batch_l, 1234)
    orc.train_step(params, adam, spec, A, x, None, noise, xl, yl)       # warm-up
    t0, n = time.perf_counter(), 0
    while True:
        orc.train_step(params, adam, spec, A, x, None, noise, xl, yl)
        n += 1
        dt = time.perf_counter() - t0
        if dt > target_s or n >= 50:
            break
    return {"value": batch * n / dt, "unit": "samples/s", "cores": threads, "kind": "port",
            "sample": f"{n} steps of the semi-supervised step at U={batch}, L={batch_l} (oracle/cdgvae_oracle.py, torch CPU fp32)",
            "ms_per_step": 1e3 * dt / n}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    batch, batch_l = 1024, 256
    from oracle import cdgvae_oracle as orc
    torch.set_num_threads(threads)
    cfg = make_config(batch, batch_l)
    spec = orc.pendulum_spec(cfg, orc.pendulum_masks(64))
    A = orc.i_b_inv(orc.pendulum_B(4))
    params = orc.init_params(spec, 1)
    adam = orc.new_adam_state(params)
    x, xl, yl, noise = synth(batch, batch_l, 1234)
    for _ in range(max(1, args.warmup)):
        orc.train_step(params, adam, spec, A, x, None, noise, xl, yl)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        orc.train_step(params, adam, spec, A, x, None, noise, xl, yl)
    dt = time.perf_counter() - t0
    v = batch * args.steps / dt
    sample = f"each step = the semi-supervised step on a bounded sample U={batch}, L={batch_l} of the workload"
    line = {"impl": "reference", "metric": "train samples/sec", "value": v, "unit": "samples/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "pendulum CDG-VAE semi-supervised (main_semi.py) training step", "sample": sample},
            "cpu_baseline": {"value": v, "unit": "samples/s", "cores": threads, "kind": "port", "sample": sample},
            "e2e": {"value": v, "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=1 << 17, help="unlabeled samples per GPU per step")
    ap.add_argument("--gemm", default="auto", choices=["auto", "simt", "tc3x", "tc1x", "bf3x"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import torch.distributed as dist
    from cdgvae_b200 import _lib
    from cdgvae_b200.data import DevicePrefetcher
    from cdgvae_b200.modules.model import CDGVAE
    from cdgvae_b200.modules.train import train_CDGVAE_semi_loaders
    from oracle import cdgvae_oracle as orc   # only for cpu_baseline and the shared synthetic B / masks

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    # stdout carries exactly ONE JSON line: anything libraries print meanwhile (NCCL's version banner at the first
    # communicator creation) goes to stderr
    sys.stdout.flush()
    saved_stdout = os.dup(1)
    os.dup2(2, 1)
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("NCCL_DEBUG", "WARN")      # keep NCCL's version banner off stdout (one JSON line only)
        dist.init_process_group("nccl", device_id=dev)
    W = max(3, args.warmup)
    B, BL = args.batch, args.batch // 4
    cfg = make_config(B, BL)
    cfg["gemm_mode"] = args.gemm
    torch.manual_seed(1)
    model = CDGVAE(orc.pendulum_B(4), orc.pendulum_masks(64), cfg, "cpu").to(dev)
    opt = torch.optim.Adam(model.parameters(), lr=cfg["lr"])

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu = cpu_baseline(os.cpu_count() or 1)

    xd, xld, yld, nd, xu8, xlu8 = synth_device(B, BL, 1234 + rank, dev)
    model.noise_fn = lambda n, d: nd

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident timing: inputs already in HBM ------------------------------------------
    loaderU, loaderL = [xd] * 1, [(xld, yld)] * 1
    for _ in range(W):
        train_CDGVAE_semi_loaders(loaderL, loaderU, model, cfg, opt, dev)
    model.profile(True)
    model.profile_read()
    sampler = ClockSampler(local)
    sampler.start()
    n0 = _lib.lib().cdg_launch_count()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    train_CDGVAE_semi_loaders([(xld, yld)] * args.steps, [xd] * args.steps, model, cfg, opt, dev)
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    launches = _lib.lib().cdg_launch_count() - n0
    sampler.stop_flag = True
    sampler.join()
    prof = model.profile_read()
    model.profile(False)
    t = torch.tensor([ms], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t)
    value = B * world * args.steps / (ms / 1e3)

    # ---- end to end: pinned host inputs, H2D every step, logs read back -------------------------
    e2e = None
    if not args.no_e2e:
        def pinned(t):
            h = torch.empty(t.shape, dtype=t.dtype, pin_memory=True)
            h.copy_(t)
            return h
        ylp, noise_p = pinned(yld), pinned(nd)
        model.noise_fn = lambda n, d: noise_p                     # CPU noise, copied H2D per step (model.py:276)

        def measure(xs, xls, pixels):
            class Marked:                # starts the clock on the compute stream when step `at` is about to be enqueued
                exact_len = True

                def __init__(self, loader, at):
                    self.loader, self.at = loader, at

                def __len__(self):
                    return len(self.loader)

                def __iter__(self):
                    for i, b in enumerate(self.loader):
                        if i == self.at:
                            e0.record()
                        yield b

            def run(k, mark_at=None):
                pu = DevicePrefetcher([xs] * k, dev, pixels=pixels)
                return train_CDGVAE_semi_loaders(DevicePrefetcher([(xls, ylp)] * k, dev, pixels=pixels),
                                                 Marked(pu, mark_at) if mark_at is not None else pu, model, cfg, opt, dev)

            def reduce_ms():
                t = torch.tensor([e0.elapsed_time(e1)], device=dev)
                if world > 1:
                    dist.all_reduce(t, op=dist.ReduceOp.MAX)
                return float(t)
            # steady state: ONE loop of W + K steps, every step with its own H2D copy; the clock starts after the W warm-up
            # steps (pipeline full, as in any epoch longer than a few steps) and stops when the logs are back on the host
            barrier()
            We = max(W, 6)               # the staging ring needs a few steps to settle into its copy-bound rhythm
            run(We + args.steps, mark_at=We)
            e1.record()
            barrier()
            warm = reduce_ms()
            # cold: K steps from an empty pipeline (the first copy cannot overlap anything)
            barrier()
            e0.record()
            run(args.steps)
            e1.record()
            barrier()
            cold = reduce_ms()
            h2d = xs.numel() * xs.element_size() + xls.numel() * xls.element_size() + ylp.numel() * 4 + noise_p.numel() * 4
            return {"value": B * world * args.steps / (warm / 1e3), "unit": "samples/s", "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": 4 * 8, "ms_per_step": warm / args.steps,
                    "timed_region": f"steps {We + 1}..{We + args.steps} of one {We + args.steps}-step loop (pipeline warm)",
                    "cold_start_ms_per_step": cold / args.steps}

        # headline: the host holds the images as the dataset stores them (uint8 pixels, modules/datasets.py:24-27); the
        # (p - 127.5) / 127.5 of datasets.py:28 runs on the device (cdg_pixels_to_float), bit-identical fp32 batches
        e2e = measure(pinned(xu8), pinned(xlu8), True)
        e2e["api"] = ("cdgvae_b200.modules.train.train_CDGVAE_semi_loaders + data.DevicePrefetcher(pixels=True): pinned host "
                      "batches of uint8 image bytes, converted on the device exactly as modules/datasets.py:28 does on the host")
        # the same with the host holding the already converted fp32 images (4 bytes per pixel over PCIe)
        e2e_fp32 = measure(pinned(xd), pinned(xld), False)
        e2e_fp32["api"] = "same call, data.DevicePrefetcher over pinned fp32 host batches"
        e2e["fp32_host_batches"] = e2e_fp32

    if rank == 0:
        hbm, tf_burst, tf_sus, src = peaks()
        gemm_ms = sum(prof[k] for k in ("enc0_fwd", "dec2_fwd", "dec2_dgrad", "dec2_wgrad", "enc0_wgrad", "gemm_other"))
        flops = 2.0 * (MACS_U * B + MACS_L * BL) * args.steps
        ach = flops / (gemm_ms / 1e3) / 1e12 if gemm_ms > 0 else 0.0
        line = {
            "metric": "train samples/sec", "value": value, "unit": "samples/s", "n_gpus": world, "steps": args.steps,
            "warmup": W, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "pendulum CDG-VAE semi-supervised (main_semi.py) training step, BASELINE configs[1]",
                       "batch_per_gpu": B, "batch_labeled_per_gpu": BL, "global_batch": B * world, "scm": "nonlinear",
                       "image": "64x64x3", "gemm_mode": args.gemm, "parallelism": f"dp{world}",
                       "l2_policy": f"inputs larger than L2: x is {B * 49152 / 1e9:.1f} GB per step",
                       "samples_counted": "unlabeled samples (the labeled quarter-batch rides along)"},
            "roofline": {"bound": "tensor", "kernel": "the GEMM kernel family (all Linear fwd/dgrad/wgrad launches of the step)",
                         "achieved": ach, "peak": tf_sus, "unit": "TFLOP/s", "frac": ach / tf_sus,
                         "traffic": NCU_TRAFFIC["bytes"], "traffic_launch": NCU_TRAFFIC["launch"],
                         "traffic_algorithmic": NCU_TRAFFIC["algorithmic"],
                         "frac_of_fp32_faithful_ceiling": ach / (tf_sus / (3.0 if args.gemm in ("auto", "bf3x") else 6.0)),
                         "peak_source": f"bf16 dense sustained, {src} (MEASURED_PEAKS.json)",
                         "algorithmic_flops_per_step": flops / args.steps, "gemm_ms_per_step": gemm_ms / args.steps,
                         "note": "fp32-faithful arithmetic (parity 1e-4): every product is 3 bf16 MMAs on hi/lo splits (bf16x3; "
                                 "gemm=tc3x: 3 TF32 MMAs = 6 bf16-equivalents), so frac <= 1/3 (1/6) by construction",
                         "breakdown_ms_per_step": {k: v / args.steps for k, v in prof.items()}},
            "cpu_baseline": cpu,
            "e2e": e2e,
            "gpu_launches": int(launches),
            "clocks": sampler.result(),
        }
        sys.stdout.flush()
        os.dup2(saved_stdout, 1)
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
