"""Synthetic inputs of SURVEY.md §8(d): the structural matrices, masks and seeded batches every
golden case, GPU test, smoke() and bench.py workload is built from.  Shared by the product-side drivers
(bench.py, __graft_entry__.py) and by the oracle / tests; it belongs to neither the product package nor
oracle/ (the reference builds these in its entry scripts, which are out of scope: main.py:137-179,
tabular/main.py:138-165, tabular/main_tvae.py:174-192)."""
from typing import List

import torch

Tensor = torch.Tensor


def pendulum_B(node: int = 4) -> Tensor:
    """main.py:137-147 with dataset.name = ['light','angle','length','position']."""
    B = torch.zeros(node, node)
    B[0, 2] = B[0, 3] = B[1, 2] = B[1, 3] = 1
    indeg = B.sum(0)
    m = indeg != 0
    B[:, m] = B[:, m] / indeg[m]
    return B


def pendulum_masks(image_size: int = 64, bands=(20, 51)) -> List[Tensor]:
    """main.py:167-179 row-band masks; `bands` scales with image_size for small test cases."""
    lo = [0, bands[0], bands[1]]
    hi = [bands[0], bands[1], image_size]
    out = []
    for a, b in zip(lo, hi):
        m = torch.zeros(image_size, image_size, 3)
        m[a:b, ...] = 1
        out.append(m)
    return out


def tabular_B(dataset: str) -> Tensor:
    """tabular/main.py:138-165."""
    if dataset in ("loan", "adult"):
        B = torch.zeros(3, 3)
        B[:-1, -1] = 1
    elif dataset == "covtype":
        B = torch.zeros(6, 6)
        B[[0, 3, 4, 5], 1] = 1
        B[[3, 4, 5], 2] = 1
        B[[0, 5], 3] = 1
    else:
        raise ValueError("Not supported dataset!")
    indeg = B.sum(0)
    m = indeg != 0
    B[:, m] = B[:, m] / indeg[m]
    return B


def synth_pendulum(batch: int, image_size: int = 64, node: int = 4, seed: int = 1234, noise_seed: int = 4321):
    g = torch.Generator().manual_seed(seed)
    x = torch.rand(batch, image_size, image_size, 3, generator=g) * 2 - 1
    white = torch.rand(batch, image_size, image_size, 3, generator=g) < 0.9
    x = torch.where(white, torch.ones_like(x), x)
    y = torch.rand(batch, node + 1, generator=g)
    noise = torch.randn(batch, node, generator=torch.Generator().manual_seed(noise_seed))
    return x, y, noise


def synth_tabular(dataset: str, batch: int, seed: int = 1234, noise_seed: int = 4321):
    g = torch.Generator().manual_seed(seed)
    if dataset in ("loan", "adult"):
        x = torch.randn(batch, 5, generator=g)
        x[:, 0] = (torch.rand(batch, generator=g) < 0.25).float()
        y = torch.rand(batch, 3, generator=g)
        d = 3
    elif dataset == "covtype":
        x = torch.randn(batch, 8, generator=g)
        x[:, 7] = torch.randint(1, 8, (batch,), generator=g).float()
        y = torch.rand(batch, 6, generator=g)
        d = 6
    else:
        raise ValueError("Not supported dataset!")
    noise = torch.randn(batch, d, generator=torch.Generator().manual_seed(noise_seed))
    return x, y, noise


def tvae_shape(kind: str):
    """SURVEY §8(d) cfg 4: loan-shaped (5 continuous x (1 tanh + 5 softmax)) and covtype-shaped."""
    if kind == "loan":
        oil = [[(1, "tanh"), (5, "softmax")] for _ in range(5)]
        mask_ = [0, 2, 2, 1]
        d, B = 3, tabular_B("loan")
    elif kind == "covtype":
        oil = [[(1, "tanh"), (5, "softmax")] for _ in range(7)] + [[(7, "softmax")]]
        mask_ = [0, 1, 1, 2, 1, 1, 2]
        d, B = 6, tabular_B("covtype")
    else:
        raise ValueError(kind)
    dims = [sum(s[0] for s in col) for col in oil]
    cs = [sum(mask_[: j + 1]) for j in range(len(mask_))]
    mask = [sum(dims[cs[j]: cs[j + 1]]) for j in range(len(mask_) - 1)]   # main_tvae.py:174-192
    return oil, mask, d, B, sum(dims)


def synth_tvae(kind: str, batch: int, seed: int = 1234, noise_seed: int = 4321):
    oil, mask, d, B, D = tvae_shape(kind)
    g = torch.Generator().manual_seed(seed)
    x = torch.zeros(batch, D)
    st = 0
    for col in oil:
        for dim, fn in col:
            if fn != "softmax":
                x[:, st] = torch.rand(batch, generator=g) * 1.98 - 0.99
            else:
                idx = torch.randint(0, dim, (batch,), generator=g)
                x[torch.arange(batch), st + idx] = 1.0
            st += dim
    y = torch.rand(batch, d, generator=g)
    noise = torch.randn(batch, d, generator=torch.Generator().manual_seed(noise_seed))
    return x, y, noise


def celeba_B(node: int = 6, structure: int = 0) -> Tensor:
    """celeba/main.py:86-108 with dataset.nodes order [Smiling, Male, High_Cheekbones, Mouth_Slightly_Open,
    Narrow_Eyes, Chubby] (celeba/module/datasets.py), adjacency scaling on."""
    B = torch.zeros(node, node)
    if structure == 0:
        for j in (2, 3, 5, 4):
            B[0, j] = 1
        B[1, 4] = 1
    indeg = B.sum(0)
    m = indeg != 0
    B[:, m] = B[:, m] / indeg[m]
    return B


def synth_celeba(batch: int, seed: int = 1234, noise_seed: int = 4321, size: int = 128):
    """SURVEY.md §8(d) cfg 5: channels 0-2 U(0,1), channels 3-7 Bernoulli(0.5) masks, y Bernoulli(0.5)."""
    g = torch.Generator().manual_seed(seed)
    img = torch.rand(batch, size, size, 3, generator=g)
    msk = (torch.rand(batch, size, size, 5, generator=g) < 0.5).float()
    y = (torch.rand(batch, 6, generator=g) < 0.5).float()
    gn = torch.Generator().manual_seed(noise_seed)
    n1 = torch.randn(batch, 6, generator=gn)
    n2 = torch.randn(batch, 6, generator=gn)
    return torch.cat([img, msk], -1), y, n1, n2
