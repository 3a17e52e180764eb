"""The reference arm of bench.py: the UNMODIFIED reference train loops, imported by file path from baseline/_ref/
(tools/install_reference.py copies them there; /root/reference does not exist on the GPU box), run on the host cores
(`device="cpu"`, the reported CPU baseline) or on the B200 through stock PyTorch (`device="cuda"`, informational: the
"existing GPU path" of BASELINE.md §3).  None of this repository's kernels, models or engine is on this path.

Every runner returns {"value": units/s, "unit", "ms_per_step", "steps", "sample", "cores"|"device"} for a bounded sample
of the workload: `units` per step, `warmup` untimed steps, then timed steps until `steps` are done or `target_s` elapsed.
"""
import importlib.util
import os
import sys
import time
from collections import namedtuple

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import synthetic_inputs as syn  # noqa: E402

REF = os.path.join(ROOT, "baseline", "_ref")
_mods = {}


def available():
    return os.path.exists(os.path.join(REF, "modules", "train.py"))


def _load(name, rel):
    if name not in _mods:
        spec = importlib.util.spec_from_file_location(name, os.path.join(REF, rel))
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        if hasattr(mod, "tqdm"):
            mod.tqdm.tqdm = lambda it, **kw: it            # progress bars off; nothing else is touched
        _mods[name] = mod
    return _mods[name]


def pendulum_modules():
    return _load("ref_pend_model", "modules/model.py"), _load("ref_pend_train", "modules/train.py")


def tabular_modules():
    return _load("ref_tab_model", "tabular/modules/model.py"), _load("ref_tab_train", "tabular/modules/train.py")


def celeba_modules():
    """celeba/module/model.py:117 asks torchvision for ImageNet weights (`pretrained=True`): there is no network and no cached
    checkpoint, so the factory is wrapped to return the randomly initialised network, as in tests/golden/make_golden_celeba.py."""
    if "celeba" not in _mods:
        sys.path.insert(0, os.path.join(REF, "celeba"))
        saved = {k: sys.modules.pop(k) for k in list(sys.modules) if k == "module" or k.startswith("module.")}
        try:
            import module.model as rm
            import module.train as rt
        finally:
            sys.path.remove(os.path.join(REF, "celeba"))
            for k in [k for k in sys.modules if k == "module" or k.startswith("module.")]:
                _mods["celeba:" + k] = sys.modules.pop(k)
            sys.modules.update(saved)
        rt.tqdm.tqdm = lambda it, **kw: it
        _mods["celeba"] = (rm, rt)
    return _mods["celeba"]


class _random_resnet18:
    """`models.resnet18(pretrained=True)` -> the same architecture, randomly initialised (no download possible here)."""

    def __enter__(self):
        import torchvision
        self.tv, self.orig = torchvision, torchvision.models.resnet18
        orig = self.orig
        torchvision.models.resnet18 = lambda pretrained=False, **kw: orig(weights=None)

    def __exit__(self, *a):
        self.tv.models.resnet18 = self.orig


def _sync(device):
    if str(device).startswith("cuda"):
        torch.cuda.synchronize()


def _timed(step, units, warmup, steps, target_s, device):
    for _ in range(max(1, warmup)):
        step()
    _sync(device)
    t0, n = time.perf_counter(), 0
    while n < steps:
        step()
        n += 1
        if target_s is not None and time.perf_counter() - t0 > target_s:
            break
    _sync(device)
    dt = time.perf_counter() - t0
    return {"value": units * n / dt, "ms_per_step": 1e3 * dt / n, "steps": n}


def _finish(r, unit, sample, device, threads):
    r.update(unit=unit, sample=sample, kind="reference")
    if str(device).startswith("cuda"):
        r["device"] = torch.cuda.get_device_name(0) + " (stock PyTorch, config['cuda']=True)"
    else:
        r["cores"] = threads
    return r


class _Images(torch.utils.data.Dataset):
    """Map-style dataset of pre-made tensors, as modules/datasets.py's classes are (one item = one image [, label])."""

    def __init__(self, x, y=None):
        self.x, self.y = x, y

    def __len__(self):
        return self.x.shape[0]

    def __getitem__(self, i):
        return self.x[i] if self.y is None else (self.x[i], self.y[i])


def pendulum_config(batch, batch_l, scm="nonlinear", cuda=False):
    cfg = dict(node=4, scm=scm, flow_num=1, inverse_loop=100, factor=[1, 1, 2], image_size=64, batch_size=batch,
               batch_sizeL=batch_l, lr=1e-3, beta=0.1, seed=1, labeled_ratio=0.1, cuda=cuda)
    cfg["lambda"] = 5.0
    return cfg


def pendulum_semi(batch=1024, device="cpu", warmup=1, steps=20, target_s=None, threads=None):
    """modules/train.py:211-282 train_CDGVAE_semi, one unlabeled batch of `batch` images + batch/4 labeled per step."""
    pm, pt = pendulum_modules()
    threads = threads or os.cpu_count() or 1
    torch.set_num_threads(threads)
    cuda = str(device).startswith("cuda")
    cfg = pendulum_config(batch, batch // 4, "nonlinear", cuda)
    torch.manual_seed(1)
    model = pm.CDGVAE(syn.pendulum_B(4), syn.pendulum_masks(64), cfg, device)
    model.train()
    opt = torch.optim.Adam(model.parameters(), lr=cfg["lr"])
    x, y, _ = syn.synth_pendulum(batch, 64, 4)
    dsU, dsL = _Images(x), _Images(x[: batch // 4].roll(1, 0), y[: batch // 4])
    step = lambda: pt.train_CDGVAE_semi(dsL, dsU, model, cfg, opt, device)
    r = _timed(step, batch, warmup, steps, target_s, device)
    return _finish(r, "samples/s", f"train_CDGVAE_semi (baseline/_ref/modules/train.py, unmodified), one step per call on a bounded "
                   f"sample U={batch}, L={batch // 4} of the workload, its own DataLoader(shuffle=True) over in-memory fp32 images",
                   device, threads)


def pendulum_b128(batch=128, device="cpu", warmup=2, steps=40, target_s=None, threads=None):
    """modules/train.py:150-209 train_CDGVAE at the reference's own batch (main.py:96), linear SCM (BASELINE configs[0])."""
    pm, pt = pendulum_modules()
    threads = threads or os.cpu_count() or 1
    torch.set_num_threads(threads)
    cuda = str(device).startswith("cuda")
    cfg = pendulum_config(batch, 0, "linear", cuda)
    torch.manual_seed(1)
    model = pm.CDGVAE(syn.pendulum_B(4), syn.pendulum_masks(64), cfg, device)
    model.train()
    opt = torch.optim.Adam(model.parameters(), lr=cfg["lr"])
    x, y, _ = syn.synth_pendulum(batch, 64, 4)
    data = [(x, y)]
    step = lambda: pt.train_CDGVAE(data, model, cfg, opt, device)
    r = _timed(step, batch, warmup, steps, target_s, device)
    return _finish(r, "samples/s", f"train_CDGVAE (unmodified), batch {batch}, linear SCM", device, threads)


def tabular_config(dataset, cuda=False):
    cfg = dict(dataset=dataset, scm="linear", flow_num=1, inverse_loop=100, lr=0.01, beta=0.01, seed=1, cuda=cuda)
    cfg["lambda"] = 10.0
    if dataset in ("loan", "adult"):
        cfg.update(node=3, factor=[1, 1, 1], input_dim=5)
        mask = [2, 2, 1] if dataset == "loan" else [1, 1, 3]
        ft = [1, 2, 3, 4, 0] if dataset == "loan" else [2, 3, 0, 1, 4]
    else:
        cfg.update(node=6, factor=[1] * 6, input_dim=8)
        mask, ft = [1, 1, 2, 1, 1, 8], None
    return cfg, mask, ft


def tabular(dataset="adult", rows=1 << 16, device="cpu", warmup=1, steps=20, target_s=None, threads=None):
    """tabular/modules/train.py:173-243 train_CDGVAE, one batch of `rows` rows per step."""
    tm, tt = tabular_modules()
    threads = threads or os.cpu_count() or 1
    torch.set_num_threads(threads)
    cuda = str(device).startswith("cuda")
    cfg, mask, ft = tabular_config(dataset, cuda)
    cfg["batch_size"] = rows
    torch.manual_seed(1)
    model = tm.CDGVAE(syn.tabular_B(dataset), mask, cfg, device)
    model.train()
    opt = torch.optim.Adam(model.parameters(), lr=cfg["lr"])
    x, y, _ = syn.synth_tabular(dataset, rows)
    ds = namedtuple("DS", ["flatten_topology"])(ft)
    data = [(x, y)]
    step = lambda: tt.train_CDGVAE(ds, data, model, cfg, opt, device)
    r = _timed(step, rows, warmup, steps, target_s, device)
    return _finish(r, "rows/s", f"tabular train_CDGVAE (unmodified), dataset={dataset}, one batch of {rows} rows per step", device, threads)


def tvae_config(kind, cuda=False):
    oil, mask, d, Bm, D = syn.tvae_shape(kind)
    cfg = dict(dataset=kind, scm="linear", flow_num=1, inverse_loop=100, lr=1e-3, weight_decay=1e-5, seed=1, node=d,
               factor=[1] * d, input_dim=D, sigma_range=[0.01, 0.1] if kind == "loan" else [0.005, 0.01], cuda=cuda)
    cfg["lambda"] = 5.0
    return cfg, oil, mask, Bm


def tvae(kind="loan", rows=1 << 15, device="cpu", warmup=1, steps=20, target_s=None, threads=None):
    """tabular/modules/train.py:245-320 train_TVAE, one batch of `rows` rows per step."""
    tm, tt = tabular_modules()
    threads = threads or os.cpu_count() or 1
    torch.set_num_threads(threads)
    cuda = str(device).startswith("cuda")
    cfg, oil, mask, Bm = tvae_config(kind, cuda)
    cfg["batch_size"] = rows
    torch.manual_seed(1)
    model = tm.TVAE(Bm, mask, cfg, device)
    model.train()
    opt = torch.optim.Adam(model.parameters(), lr=cfg["lr"], weight_decay=cfg["weight_decay"])
    Span = namedtuple("SpanInfo", ["dim", "activation_fn"])
    ref_oil = [[Span(*s) for s in col] for col in oil]
    x, y, _ = syn.synth_tvae(kind, rows)
    data = [(x, y)]
    step = lambda: tt.train_TVAE(ref_oil, None, data, model, cfg, opt, device)
    r = _timed(step, rows, warmup, steps, target_s, device)
    return _finish(r, "rows/s", f"train_TVAE (unmodified), {kind}-shaped table, one batch of {rows} rows per step", device, threads)


def celeba_config(batch, cuda=False):
    cfg = dict(node=6, latent_dim=6, scm="linear", flow_num=1, inverse_loop=100, beta=0.1, lr=1e-3, seed=1,
               batch_size=batch, cuda=cuda, pretrained=False)
    cfg["lambda"] = 5.0
    return cfg


def celeba(batch=16, device="cpu", warmup=1, steps=2, target_s=None, threads=None):
    """celeba/module/train.py:10-76 train_CDGVAE at the reference's batch 16 (celeba/main.py:70), random-init frozen ResNet-18."""
    rm, rt = celeba_modules()
    threads = threads or os.cpu_count() or 1
    torch.set_num_threads(threads)
    cuda = str(device).startswith("cuda")
    cfg = celeba_config(batch, cuda)
    x, y, _, _ = syn.synth_celeba(batch)
    if cuda:
        x, y = x.cuda(), y.cuda()
    masks = torch.split(x[..., 3:], 1, dim=-1)
    torch.manual_seed(1)
    with _random_resnet18():
        model = rm.CDGVAE(syn.celeba_B(), masks, cfg, device)
    if cuda:
        model = model.cuda()
    model.train()
    opt = torch.optim.Adam(model.parameters(), lr=cfg["lr"])
    data = [(x, y)]
    step = lambda: rt.train_CDGVAE(data, model, cfg, opt, device)
    r = _timed(step, batch, warmup, steps, target_s, device)
    return _finish(r, "samples/s", f"celeba train_CDGVAE (unmodified; resnet18 weights=None), batch {batch}, 128x128", device, threads)


if __name__ == "__main__":
    import json
    which = sys.argv[1] if len(sys.argv) > 1 else "pendulum_semi"
    dev = sys.argv[2] if len(sys.argv) > 2 else "cpu"
    fn = {"pendulum_semi": pendulum_semi, "pendulum_b128": pendulum_b128, "tabular": tabular, "tvae": tvae, "celeba": celeba}[which]
    print(json.dumps(fn(device=dev, steps=3)))
