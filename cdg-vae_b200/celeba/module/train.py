"""Drop-in for the reference's celeba/module/train.py.

    train_CDGVAE(train_loader, model, config, optimizer, device) -> (logs, xhat)        celeba/module/train.py:10-76

Per batch the reference does H2D, zero_grad, forward, L1 reconstruction + two KL terms + label alignment, backward,
Adam and five `.item()` syncs; here a batch is two C-ABI calls (cdg_celeba_step, cdg_adam_step), the log row stays on
the device and is read back once per call.
"""
import torch

from ... import dist as _dist
from ...modules.train import _lookahead
from .model import LOG_KEYS


GRAPH_MAX_BATCH = 64


def _step(model, x_batch, y_batch, noise, row, last):
    """One training step.  At the reference's batch sizes the step is ~780 launches on six streams and the HOST sets the pace:
    the generators' chains are enqueued one after the other (~1.2 ms of launch calls each), so the fifth generator's stream
    starts ~5 ms after the first although they could run side by side.  Replayed as a CUDA graph (the internal streams become
    parallel branches of it) the step drops from 13.0 to ~9 ms at batch 16.  Larger batches are GPU-bound and run eagerly."""
    from ...engine import _f32c
    dp = _dist.world() > 1
    if (x_batch.shape[0] > GRAPH_MAX_BATCH or not getattr(model, "use_graphs", True) or
            (dp and not getattr(model, "dp_graphs", True))):
        out = model.forward_backward(x_batch, y_batch, noise, row, xhat=last)
        model.adam_step(grad_scale=model.exchange_gradients())
        return out["xhat"] if last else None
    dev = model.arena_device
    ins = {"x": _f32c(x_batch, dev), "y": _f32c(y_batch, dev), "n1": _f32c(noise[0], dev), "n2": _f32c(noise[1], dev)}
    g = model._opt_group
    lr = g["lr"]
    key = ("celeba", tuple(tuple(v.shape) for v in ins.values()), float(lr.item() if torch.is_tensor(lr) else lr),
           tuple(g["betas"]), g["eps"], g["weight_decay"], model.config.get("beta"), model.config.get("lambda"), _dist.world())

    def body(t):
        out_row = torch.empty_like(row)
        # the reconstruction is written by every replay (3 MB at batch 16): one graph whether or not the caller wants it
        out = model.forward_backward(t["x"], t["y"], (t["n1"], t["n2"]), out_row, xhat=True)
        model.adam_step(grad_scale=model.exchange_gradients() if dp else 1.0)
        return {"row": out_row, "xhat": out["xhat"]}

    try:
        outs = model.graphed_step(key, ins, body)
    except RuntimeError as e:                         # something in the step cannot be captured here: eager from now on
        import warnings
        warnings.warn(f"CelebA step: CUDA-graph capture failed ({e}); falling back to eager steps")
        model.use_graphs = False
        model.drop_graphs()
        out = model.forward_backward(x_batch, y_batch, noise, row, xhat=last)
        model.adam_step(grad_scale=model.exchange_gradients())
        return out["xhat"] if last else None
    row.copy_(outs["row"], non_blocking=True)
    return outs["xhat"].clone() if last else None


def train_CDGVAE(train_loader, model, config, optimizer, device):
    for k in ("beta", "lambda"):                     # the loop reads beta / lambda from `config` (train.py:65-66)
        if k in config:
            model.config[k] = config[k]
    model.bind_optimizer(optimizer)
    xhat, n = None, 0
    # sized loaders are counted, anything else is read one batch ahead, to know which batch is the last (whose xhat is returned)
    for (x_batch, y_batch), last in _lookahead(train_loader):
        rows = model._log_rows(n + 1, len(LOG_KEYS))
        noise = (model._noise(x_batch.shape[0]), model._noise(x_batch.shape[0]))        # model.py:182, :184
        out = _step(model, x_batch, y_batch, noise, rows[n], last)
        if last:
            xhat = out                                # the reference returns the last batch's reconstruction (train.py:76)
        n += 1
    logs = {k: [] for k in LOG_KEYS}
    if n:
        rows = model._logs[:n]
        if _dist.world() > 1:
            rows = _dist.allreduce_mean_(rows.clone())
        host = rows.cpu()
        for j, k in enumerate(LOG_KEYS):
            logs[k] = host[:, j].tolist()
    model._grad_views()
    return logs, xhat
