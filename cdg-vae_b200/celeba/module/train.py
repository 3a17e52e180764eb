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
        out = model.forward_backward(x_batch, y_batch, noise, rows[n], xhat=last)
        model.adam_step(grad_scale=model.exchange_gradients())
        if last:
            xhat = out["xhat"]                        # the reference returns the last batch's reconstruction (train.py:76)
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
