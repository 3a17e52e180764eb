"""Throughput of the tabular CDG-VAE / CDG-TVAE step (BASELINE configs[2], [3]) in rows/s on one B200, with the
CPU oracle timed beside it on a bounded sample.  Prints one JSON line per family."""
import json
import os
import sys
import time
from collections import namedtuple

import torch

sys.path.insert(0, ".")
from cdgvae_b200.tabular.modules import model as M, train as T  # noqa: E402
from oracle import cdgvae_oracle as orc  # noqa: E402

ROWS = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
if os.environ.get("CDG_TVAE_ROUTE"):                       # 1 = FFMA tiles (default), 2 = mma.sync fragments, 0 = one row per thread
    from cdgvae_b200 import _lib as _L
    _L.lib().cdg_tabular_tvae_tile(int(os.environ["CDG_TVAE_ROUTE"]))
STEPS = 10
DS = namedtuple("DS", ["flatten_topology"])
Span = namedtuple("SpanInfo", ["dim", "activation_fn"])


def run(family, name):
    if family == "tabular":
        d = 6 if name == "covtype" else 3
        cfg = dict(dataset=name, scm="linear", flow_num=1, inverse_loop=100, lr=0.01, beta=0.01, node=d, factor=[1] * d,
                   input_dim=8 if name == "covtype" else 5)
        cfg["lambda"] = 10.0
        mask = {"loan": [2, 2, 1], "adult": [1, 1, 3], "covtype": [1, 1, 2, 1, 1, 8]}[name]
        ft = {"loan": [1, 2, 3, 4, 0], "adult": [2, 3, 0, 1, 4], "covtype": None}[name]
        Bm = orc.tabular_B(name)
        spec = orc.tabular_spec(cfg, mask, ft)
        torch.manual_seed(1)
        model = M.CDGVAE(Bm, mask, cfg, "cpu").to("cuda")
        opt = torch.optim.Adam(model.parameters(), lr=cfg["lr"])
        x, y, noise = orc.synth_tabular(name, ROWS)
        step = lambda data: T.train_CDGVAE(DS(ft), data, model, cfg, opt, "cuda")
        bytes_row = 4 * (x.shape[1] + 2 * d)
    else:
        oil, mask, d, Bm, D = orc.tvae_shape(name)
        cfg = dict(dataset=name, scm="linear", flow_num=1, inverse_loop=100, lr=1e-3, weight_decay=1e-5, node=d,
                   factor=[1] * d, input_dim=D, sigma_range=[0.01, 0.1] if name == "loan" else [0.005, 0.01])
        cfg["lambda"] = 5.0
        spec = orc.tvae_spec(cfg, mask, oil)
        torch.manual_seed(1)
        model = M.TVAE(Bm, mask, cfg, "cpu").to("cuda")
        opt = torch.optim.Adam(model.parameters(), lr=cfg["lr"], weight_decay=cfg["weight_decay"])
        x, y, noise = orc.synth_tvae(name, ROWS)
        o = [[Span(*s) for s in col] for col in oil]
        step = lambda data: T.train_TVAE(o, None, data, model, cfg, opt, "cuda")
        bytes_row = 4 * (D + 2 * d)
    xd, yd, nd = x.cuda(), y.cuda(), noise.cuda()
    model.noise_fn = lambda n, dd: nd
    step([(xd, yd)] * 6)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    logs = step([(xd, yd)] * STEPS)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / STEPS
    # CPU oracle on a bounded sample
    torch.set_num_threads(os.cpu_count())
    n_cpu = 4096
    params = orc.init_params(spec, 1)
    adam = orc.new_adam_state(params)
    A = orc.i_b_inv(Bm)
    orc.train_step(params, adam, spec, A, x[:n_cpu], y[:n_cpu], noise[:n_cpu])
    t0 = time.perf_counter()
    k = 0
    while time.perf_counter() - t0 < 3.0:
        orc.train_step(params, adam, spec, A, x[:n_cpu], y[:n_cpu], noise[:n_cpu])
        k += 1
    cpu_rows = n_cpu * k / (time.perf_counter() - t0)
    rows_s = ROWS / (ms / 1e3)
    print(json.dumps({"family": family, "dataset": name, "rows_per_step": ROWS, "ms_per_step": ms, "rows_per_s": rows_s,
                      "hbm_bytes_per_row": bytes_row, "achieved_gbs": rows_s * bytes_row / 1e9, "loss": logs["loss"][-1],
                      "cpu_oracle_rows_per_s": cpu_rows, "cpu_cores": os.cpu_count()}), flush=True)


import builtins
_out = open("gpurun_out/tabular_rows.jsonl", "w") if os.path.isdir("gpurun_out") else None
_print = builtins.print


def print(*a, **k):                                    # tee the JSON lines into gpurun_out/
    _print(*a, **k)
    if _out:
        _out.write(" ".join(map(str, a)) + "\n")
        _out.flush()


for fam, nm in (("tabular", "loan"), ("tabular", "adult"), ("tabular", "covtype"), ("tvae", "loan"), ("tvae", "covtype")):
    run(fam, nm)
