"""Drop-in for the reference's modules/train.py: same function names, signatures and returns.

    train_CDGVAE(dataloader, model, config, optimizer, device) -> (logs, xhat)        train.py:150-209
    train_CDGVAE_semi(datasetL, datasetU, model, config, optimizer, device) -> (logs, xhat)   :211-282

Per batch the reference does H2D, zero_grad, forward, three losses, backward, Adam and 4+d `.item()`
syncs; here a batch is two C-ABI calls (forward_backward, adam_step), the log row stays on the
device and is read back once per call of these functions.
"""
import torch
from torch.utils.data import DataLoader

from .. import dist as _dist


def _log_keys(config):
    return ["loss", "recon", "KL", "alignment"] + [f"posterior_variance{i + 1}" for i in range(config["node"])]


def _lookahead(it):
    """Yield (item, is_last) so that xhat is materialised for the last batch only (train.py:209).

    Sized loaders are counted instead of read one batch ahead: pulling batch i+1 out of a DevicePrefetcher before step i is
    enqueued would put the wait for copy i+1 in front of step i on the compute stream (one whole copy of extra latency
    per pipeline fill)."""
    n = None
    if isinstance(it, (list, tuple)) or getattr(it, "exact_len", False):     # loaders whose len() is the batch count
        try:
            n = len(it)
        except TypeError:
            n = None
    if n is not None:
        i = -1
        for i, cur in enumerate(it):
            yield cur, i == n - 1
        return
    it = iter(it)
    try:
        cur = next(it)
    except StopIteration:
        return
    for nxt in it:
        yield cur, False
        cur = nxt
    yield cur, True


def _finish(model, config, n_steps):
    """One device->host copy per call instead of 4+d `.item()` per step (train.py:206-207)."""
    keys = _log_keys(config)
    logs = {k: [] for k in keys}
    if n_steps:
        rows = model._logs[:n_steps]
        if _dist.world() > 1:
            rows = _dist.allreduce_mean_(rows.clone())
        host = rows.cpu()
        for j, k in enumerate(keys):
            logs[k] = host[:, j].tolist()
    model._grad_views()
    return logs


def _hyper_key(model):
    g = model._opt_group
    lr = g["lr"]
    return (float(lr.item() if torch.is_tensor(lr) else lr), tuple(g["betas"]), g["eps"], g["weight_decay"],
            model.config.get("beta"), model.config.get("lambda"), model.gemm_mode)


def _step(model, config, optimizer, row, x, y, noise, x_l=None, y_l=None, xhat=None):
    rows = x.shape[0]
    if rows > model.GRAPH_MAX_ROWS or _dist.world() > 1 or not getattr(model, "use_graphs", True):
        model.forward_backward(x, y, noise, row, x_l=x_l, y_l=y_l, xhat=xhat)
        model.adam_step(grad_scale=model.exchange_gradients())
        return
    # launch-bound regime: replay the whole step (forward, backward, Adam) as one CUDA graph
    from ..engine import _f32c
    dev = model.arena_device
    ins = {"x": _f32c(x, dev).reshape(rows, -1), "noise": _f32c(noise, dev),
           "y": None if y is None else _f32c(y, dev), "x_l": None if x_l is None else _f32c(x_l, dev).reshape(x_l.shape[0], -1),
           "y_l": None if y_l is None else _f32c(y_l, dev)}
    key = ("pend", tuple((k, None if v is None else tuple(v.shape)) for k, v in ins.items()), xhat is not None, _hyper_key(model))

    def body(t):
        out_row = torch.empty_like(row)
        out_xhat = None if xhat is None else torch.empty_like(xhat)
        model.forward_backward(t["x"], t["y"], t["noise"], out_row, x_l=t["x_l"], y_l=t["y_l"], xhat=out_xhat)
        model.adam_step(grad_scale=1.0)
        return {"row": out_row, "xhat": out_xhat}

    outs = model.graphed_step(key, ins, body)
    row.copy_(outs["row"], non_blocking=True)
    if xhat is not None:
        xhat.copy_(outs["xhat"], non_blocking=True)


def _sync_config(model, config):
    # the train loops read beta / lambda from the `config` they are given (train.py:198-199)
    for k in ("beta", "lambda"):
        if k in config:
            model.config[k] = config[k]


def train_CDGVAE(dataloader, model, config, optimizer, device):
    _sync_config(model, config)
    model.bind_optimizer(optimizer)
    width = 4 + config["node"]
    s = config["image_size"] if "image_size" in config else model.config["image_size"]
    xhat, n = None, 0
    for (x_batch, y_batch), last in _lookahead(dataloader):
        rows = model._log_rows(n + 1, width)
        noise = model._noise(x_batch.shape[0])
        if last:
            xhat = torch.empty(x_batch.shape[0], 3 * s * s, device=model.arena_device)
        _step(model, config, optimizer, rows[n], x_batch, y_batch, noise, xhat=xhat if last else None)
        n += 1
    logs = _finish(model, config, n)
    return logs, (None if xhat is None else xhat.view(-1, s, s, 3))


def train_VAE(dataloader, model, config, optimizer, device):
    """modules/train.py:10-69: the VAE baseline's loop has the CDG-VAE losses (recon + beta KL + lambda align)."""
    return train_CDGVAE(dataloader, model, config, optimizer, device)


def train_InfoMax(dataloader, model, discriminator, config, optimizer, optimizer_D, device):
    """modules/train.py:71-148: the VAE step plus the mutual-information critic; logs gain 'MutualInfo'.  Both backward
    passes of the reference (`loss.backward(retain_graph=True); MI.backward()`) and both optimizer steps happen here."""
    _sync_config(model, config)
    model.bind_optimizer(optimizer)
    discriminator.bind_optimizer(optimizer_D)
    keys = ["loss", "recon", "KL", "alignment", "MutualInfo"] + [f"posterior_variance{i + 1}" for i in range(config["node"])]
    s = config["image_size"] if "image_size" in config else model.config["image_size"]
    perm_fn = getattr(model, "perm_fn", None)
    xhat, n = None, 0
    for (x_batch, y_batch), last in _lookahead(dataloader):
        rows = model._log_rows(n + 1, len(keys))
        noise = model._noise(x_batch.shape[0])
        perm = perm_fn(x_batch.shape[0]) if perm_fn else torch.randperm(x_batch.shape[0])      # train.py:75 (CPU draw)
        if last:
            xhat = torch.empty(x_batch.shape[0], 3 * s * s, device=model.arena_device)
        model.forward_backward(x_batch, y_batch, noise, rows[n], xhat=xhat if last else None,
                               infomax=(discriminator, perm, config["gamma"]))
        scale = model.exchange_gradients()
        discriminator.exchange_gradients()
        model.adam_step(grad_scale=scale)
        discriminator.adam_step(grad_scale=scale)
        n += 1
    logs = {k: [] for k in keys}
    if n:
        rows_ = model._logs[:n]
        if _dist.world() > 1:
            rows_ = _dist.allreduce_mean_(rows_.clone())
        host = rows_.cpu()
        for j, k in enumerate(keys):
            logs[k] = host[:, j].tolist()
    model._grad_views()
    discriminator._grad_views()
    return logs, (None if xhat is None else xhat.view(-1, s, s, 3))


def _resident_loader(model, dataset, batch_size):
    """The reference builds `DataLoader(dataset, batch_size, shuffle=True)` on every call (train.py:222-223) and pays
    per-item float64->float32 conversion + collate + a pageable H2D copy per batch.  Datasets that expose their arrays
    (`x_data` [, `y_data`], as modules/datasets.py does) are uploaded once and served by DeviceDataLoader, which yields
    the same batches in the same order under the same RNG state; anything else falls back to torch's DataLoader."""
    from ..data import DeviceDataLoader
    if not hasattr(dataset, "x_data"):
        return DataLoader(dataset, batch_size=batch_size, shuffle=True)
    cache = model.__dict__.setdefault("_resident", {})
    key = (id(dataset), int(batch_size))
    if key not in cache:
        cache[key] = DeviceDataLoader.from_dataset(dataset, batch_size, shuffle=True, device=model.arena_device)
    return cache[key]


def train_CDGVAE_semi(datasetL, datasetU, model, config, optimizer, device):
    dataloaderU = _resident_loader(model, datasetU, config["batch_size"])
    dataloaderL = _resident_loader(model, datasetL, config["batch_sizeL"])
    return train_CDGVAE_semi_loaders(dataloaderL, dataloaderU, model, config, optimizer, device)


def train_CDGVAE_semi_loaders(dataloaderL, dataloaderU, model, config, optimizer, device):
    """train_CDGVAE_semi with the two loaders supplied by the caller (any iterables of batches): the body
    of the reference loop, train.py:225-282, without the per-call DataLoader construction."""
    _sync_config(model, config)
    model.bind_optimizer(optimizer)
    width = 4 + config["node"]
    s = config["image_size"] if "image_size" in config else model.config["image_size"]
    iterL = None
    xhat, n = None, 0
    for x_batchU, last in _lookahead(dataloaderU):
        # labeled iterator restarts when exhausted (train.py:226-230)
        try:
            if iterL is None:
                raise StopIteration
            x_batchL, y_batchL = next(iterL)
        except StopIteration:
            iterL = iter(dataloaderL)
            x_batchL, y_batchL = next(iterL)
        rows = model._log_rows(n + 1, width)
        noise = model._noise(x_batchU.shape[0])
        if last:
            xhat = torch.empty(x_batchU.shape[0], 3 * s * s, device=model.arena_device)
        _step(model, config, optimizer, rows[n], x_batchU, None, noise, x_l=x_batchL, y_l=y_batchL,
              xhat=xhat if last else None)
        n += 1
    logs = _finish(model, config, n)
    return logs, (None if xhat is None else xhat.view(-1, s, s, 3))
