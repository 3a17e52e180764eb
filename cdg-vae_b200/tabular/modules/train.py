"""Drop-in for the reference's tabular/modules/train.py.

    train_VAE(dataset, dataloader, model, config, optimizer, device) -> logs             train.py:10-82
    train_CDGVAE(dataset, dataloader, model, config, optimizer, device) -> logs          train.py:173-243
    train_TVAE(output_info_list, dataset, dataloader, model, config, optimizer, device) -> logs   :245-320
    train_InfoMax(...)                                                                    train.py:84-171  (not built: raises)

One fused kernel per step (forward + losses + backward), one Adam kernel; the log rows stay on the
device until the loader is exhausted.
"""
from ... import dist as _dist


def _log_keys(config):
    return ["loss", "recon", "KL", "alignment"] + [f"posterior_variance{i + 1}" for i in range(config["node"])]


def _finish(model, config, n_steps):
    keys = _log_keys(config)
    logs = {k: [] for k in keys}
    if n_steps:
        rows = model._logs[:n_steps]
        if _dist.world() > 1:
            rows = _dist.allreduce_mean_(rows.clone())
        host = rows.cpu()
        for j, k in enumerate(keys):
            logs[k] = host[:, j].tolist()
    model._grad_views(set(model.live_param_names()))
    return logs


def _graph_step(model, x_batch, y_batch, noise, row, ft, oil, clamp):
    """One step; replayed as a CUDA graph when the batch is small enough to be launch-bound."""
    import torch
    from ...engine import _f32c
    rows = x_batch.shape[0]
    # data parallel: the three-launch step plus a 12 KB all-reduce is pure host latency when run eagerly (measured at 2 GPUs,
    # 2^20 rows: 0.29 ms per step against 0.17 on one GPU), so the NCCL call is captured into the step's graph as well
    # (`dp_graphs`, on by default; a capture failure falls back to the eager path for good)
    dp = _dist.world() > 1
    if rows > model.GRAPH_MAX_ROWS or (dp and not getattr(model, "dp_graphs", True)) or not getattr(model, "use_graphs", True):
        model.forward_backward(x_batch, y_batch, noise, row, flatten_topology=ft, output_info_list=oil)
        model.adam_step(grad_scale=model.exchange_gradients(), clamp=clamp)
        return
    dev = model.arena_device
    ins = {"x": _f32c(x_batch, dev).reshape(rows, -1), "y": _f32c(y_batch, dev), "noise": _f32c(noise, dev)}
    g = model._opt_group
    lr = g["lr"]
    key = ("tab", tuple(tuple(v.shape) for v in ins.values()), float(lr.item() if torch.is_tensor(lr) else lr),
           tuple(g["betas"]), g["eps"], g["weight_decay"], model.config.get("beta"), model.config.get("lambda"),
           None if ft is None else tuple(ft), None if oil is None else tuple(map(tuple, map(lambda c: tuple(map(tuple, c)), oil))), clamp)

    def body(t):
        out_row = torch.empty_like(row)
        model.forward_backward(t["x"], t["y"], t["noise"], out_row, flatten_topology=ft, output_info_list=oil)
        model.adam_step(grad_scale=model.exchange_gradients() if dp else 1.0, clamp=clamp)
        return {"row": out_row}

    if dp:
        try:
            outs = model.graphed_step(key + ("dp", _dist.world()), ins, body)
        except RuntimeError as e:                  # NCCL inside a capture is not available here: eager from now on
            import warnings
            warnings.warn(f"data-parallel CUDA-graph capture failed ({e}); falling back to eager steps")
            model.dp_graphs = False
            model.drop_graphs()
            model.forward_backward(x_batch, y_batch, noise, row, flatten_topology=ft, output_info_list=oil)
            model.adam_step(grad_scale=model.exchange_gradients(), clamp=clamp)
            return
    else:
        outs = model.graphed_step(key, ins, body)
    row.copy_(outs["row"], non_blocking=True)


def _sync_config(model, config):
    for k in ("beta", "lambda", "dataset", "sigma_range"):
        if k in config:
            model.config[k] = config[k]


def train_CDGVAE(dataset, dataloader, model, config, optimizer, device):
    if config["dataset"] not in ("loan", "adult", "covtype"):
        raise ValueError("Not supported dataset!")                       # train.py:210
    _sync_config(model, config)
    model.bind_optimizer(optimizer)
    ft = None
    if config["dataset"] in ("loan", "adult"):
        ft = tuple(int(v) for v in dataset.flatten_topology)             # train.py:200-202
    model._last_aux = (ft, None)
    width, n = 4 + config["node"], 0
    for (x_batch, y_batch) in iter(dataloader):
        rows = model._log_rows(n + 1, width)
        noise = model._noise(x_batch.shape[0])
        _graph_step(model, x_batch, y_batch, noise, rows[n], ft, None, None)
        n += 1
    return _finish(model, config, n)


def train_VAE(dataset, dataloader, model, config, optimizer, device):
    """tabular/modules/train.py:10-82: the loss of train_CDGVAE on the single-decoder baseline (same reconstruction terms per
    dataset, KL, alignment on all of y, beta / lambda weights), so the same step kernel runs it."""
    return train_CDGVAE(dataset, dataloader, model, config, optimizer, device)


def train_InfoMax(dataset, dataloader, model, discriminator, config, optimizer, optimizer_D, device):
    """tabular/modules/train.py:84-171.  Not built: the mutual-information term couples every row with a permuted row of
    the same batch (permute_dims, :86-90), which the one-row-per-thread step kernel cannot express in its single pass.  The
    pendulum InfoMax baseline (modules/train.py::train_InfoMax) is built; this one fails loudly instead of silently running
    the reference's eager path."""
    raise NotImplementedError("tabular train_InfoMax is not provided by cdgvae_b200 (DESIGN.md section 7); "
                              "use the reference implementation for this baseline")


def train_TVAE(output_info_list, dataset, dataloader, model, config, optimizer, device):
    _sync_config(model, config)
    model.bind_optimizer(optimizer)
    oil = [[(int(s.dim), str(s.activation_fn)) if hasattr(s, "dim") else (int(s[0]), str(s[1])) for s in col]
           for col in output_info_list]
    model._last_aux = (None, oil)
    lo, hi = config["sigma_range"]
    sig = model._offsets["sigma"]
    D = model.config["input_dim"]
    width, n = 4 + config["node"], 0
    for (x_batch, y_batch) in iter(dataloader):
        rows = model._log_rows(n + 1, width)
        noise = model._noise(x_batch.shape[0])
        # optimizer.step() then sigma.data.clamp_(lo, hi)  (train.py:313-314)
        _graph_step(model, x_batch, y_batch, noise, rows[n], None, oil, (sig, D, float(lo), float(hi)))
        n += 1
    return _finish(model, config, n)
