"""Host-side logic that needs no GPU: the C-ABI library loads and exports every declared symbol,
the drop-in model reproduces the reference's same-seed initialisation / state_dict layout, mask and
segment bookkeeping is exact, and loading the product without CUDA fails loudly."""
import ctypes
import os
import re

import pytest
import torch

import cdgvae_b200
from cdgvae_b200 import _lib
from cdgvae_b200.modules.model import CDGVAE, mask_ranges
from oracle import cdgvae_oracle as orc
from helpers import case_setup, exact_check

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    L = _lib.lib()
    header = open(os.path.join(ROOT, "include", "cdgvae.h")).read()
    declared = set(re.findall(r"\b(cdg_[a-z_0-9]+)\s*\(", header))
    assert declared == set(_lib.EXPORTS), declared ^ set(_lib.EXPORTS)
    for name in declared:
        assert hasattr(L, name), name
    assert L.cdg_version() >= 100


def test_struct_sizes_match_header():
    """ctypes mirrors vs the C compiler's view of include/cdgvae.h."""
    import subprocess, tempfile
    src = '#include "cdgvae.h"\n#include <stdio.h>\nint main(){printf("%zu %zu %zu %zu %zu %zu %zu %zu %zu %zu %zu %zu %zu %zu %zu %zu\\n",' \
          'sizeof(cdg_linear),sizeof(cdg_adam_args),sizeof(cdg_pendulum_config),sizeof(cdg_pendulum_io),' \
          'sizeof(cdg_pendulum_fwd_io),sizeof(cdg_tabular_config),sizeof(cdg_tabular_io),sizeof(cdg_conv),sizeof(cdg_bnorm),' \
          'sizeof(cdg_gen_block),sizeof(cdg_generator),sizeof(cdg_res_block),sizeof(cdg_celeba_config),sizeof(cdg_celeba_io),' \
          'sizeof(cdg_tvae_column),sizeof(cdg_tvae_transform_config));return 0;}'
    with tempfile.TemporaryDirectory() as td:
        c = os.path.join(td, "s.c")
        open(c, "w").write(src)
        exe = os.path.join(td, "s")
        subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), c, "-o", exe])
        sizes = list(map(int, subprocess.check_output([exe]).split()))
    mine = [ctypes.sizeof(t) for t in (_lib.Linear, _lib.AdamArgs, _lib.PendulumConfig, _lib.PendulumIO,
                                       _lib.PendulumFwdIO, _lib.TabularConfig, _lib.TabularIO, _lib.Conv, _lib.BNorm, _lib.GenBlock,
                                       _lib.GeneratorDesc, _lib.ResBlock, _lib.CelebaConfig, _lib.CelebaIO, _lib.TvaeColumn,
                                       _lib.TvaeTransformConfig)]
    assert sizes == mine


def _pendulum_model(c, device="cpu"):
    spec, Bm, batches, cfg = case_setup(c)
    torch.manual_seed(cfg["seed"])
    model = CDGVAE(Bm, spec.mask, cfg, device)
    return model, spec, Bm, batches, cfg


@pytest.mark.parametrize("name", ["pendulum_small_linear", "pendulum_small_nonlinear", "pendulum_full_linear"])
def test_pendulum_same_seed_init_and_state_dict(golden, name):
    c = golden(name)
    model, spec, Bm, batches, cfg = _pendulum_model(c)
    sd = model.state_dict()
    assert list(sd) == list(c["init"])                       # key names AND order
    for k, v in sd.items():
        exact_check(v, c["init"][k], k)                       # bit-exact same-seed init
    assert model.I_B_inv.tolist() == c["I_B_inv"]
    # parameters are views of one arena
    base = model._arena.data_ptr()
    for n, p in model.named_parameters():
        assert p.data_ptr() == base + 4 * model._offsets[n]
        assert model._offsets[n] % 4 == 0 or n.startswith("flows.")
    # load_state_dict writes through to the arena
    sd2 = {k: torch.full_like(v, 0.5) for k, v in sd.items()}
    model.load_state_dict(sd2)
    assert float(model._arena[model._offsets["encoder.2.bias"]]) == 0.5


def test_mask_ranges_bit_exact():
    m = orc.pendulum_masks(64, (20, 51))
    assert mask_ranges(m, 12288) == [(0, 3840), (3840, 9792), (9792, 12288)]
    bad = torch.zeros(64, 64, 3)
    bad[::2] = 1
    with pytest.raises(ValueError):
        mask_ranges([bad], 12288)              # not a band: the model switches to the general-mask path


def test_adam_segments_skip_dead_decoder_rows(golden):
    model, *_ = _pendulum_model(golden("pendulum_full_linear"))
    segs = model.adam_segments()
    live = sum(n for _, n in segs)
    total = sum(p.numel() for p in model.parameters())
    assert total == 15_148_480
    dead = 2 * 12288 * 300 + 2 * 12288                         # two thirds of the decoder output rows
    assert total - dead <= live <= total - dead + 32 * 28            # + alignment padding between merged params
    assert len(segs) <= _lib.MAX_SEG


def test_asserts_and_errors_match_reference():
    cfg = dict(node=4, scm="cubic", flow_num=1, inverse_loop=100, factor=[1, 1, 2], image_size=8)
    with pytest.raises(ValueError, match="Not supported SCM!"):
        CDGVAE(orc.pendulum_B(4), orc.pendulum_masks(8, (3, 6)), cfg, "cpu")
    cfg["scm"] = "linear"
    cfg["factor"] = [1, 1, 1]
    with pytest.raises(AssertionError):
        CDGVAE(orc.pendulum_B(4), orc.pendulum_masks(8, (3, 6)), cfg, "cpu")


def test_no_cpu_fallback(golden):
    """Without a CUDA device the product path must refuse to compute."""
    model, spec, Bm, batches, cfg = _pendulum_model(golden("pendulum_small_linear"))
    if torch.cuda.is_available():
        pytest.skip("CPU-only check")
    with pytest.raises(RuntimeError, match="no CPU path"):
        model(batches[0]["x"])


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "cdg-vae_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dp, f)).read()
                assert "oracle" not in txt.replace("restatement", ""), os.path.join(dp, f)


@pytest.mark.parametrize("name", ["tabular_loan", "tabular_adult", "tabular_covtype", "tvae_loan", "tvae_covtype"])
def test_tabular_same_seed_init_and_state_dict(golden, name):
    from cdgvae_b200.tabular.modules import model as M
    c = golden(name)
    spec, Bm, batches, cfg = case_setup(c)
    torch.manual_seed(cfg["seed"])
    cls = M.TVAE if c["family"] == "tvae" else M.CDGVAE
    model = cls(Bm, c["mask"], cfg, "cpu")
    sd = model.state_dict()
    assert list(sd) == list(c["init"])
    for k, v in sd.items():
        exact_check(v, c["init"][k], k)
    assert model.I_B_inv.tolist() == c["I_B_inv"]                 # covtype: general (non-triangular) inverse
    if name == "tabular_covtype":
        live = model.live_param_names()
        assert not any(n.startswith("decoder.6.") for n in live) and "decoder.6.6.weight" in sd
        assert sum(p.numel() for p in model.parameters()) == 390
    if name in ("tabular_loan", "tabular_adult"):
        assert sum(p.numel() for p in model.parameters()) == 87


def test_device_dataloader_reproduces_dataloader_order_and_rng_stream():
    """DeviceDataLoader (here on CPU tensors) yields exactly the batches of torch's DataLoader(shuffle=True) over
    the reference's map-style datasets, and leaves the global RNG in the same state (the noise draws that follow
    in model.encode must not shift)."""
    import numpy as np
    from torch.utils.data import DataLoader, Dataset
    from cdgvae_b200.data import DeviceDataLoader

    class Labeled(Dataset):                       # shape of modules/datasets.py::LabeledDataset
        def __init__(self):
            rng = np.random.default_rng(0)
            self.x_data = rng.standard_normal((37, 4, 4, 3))
            self.y_data = rng.random((37, 5))
        def __len__(self):
            return len(self.x_data)
        def __getitem__(self, i):
            return torch.FloatTensor(self.x_data[i]), torch.FloatTensor(self.y_data[i])

    class Unlabeled(Labeled):
        def __getitem__(self, i):
            return torch.FloatTensor(self.x_data[i])

    for cls, bs, drop in ((Labeled, 8, False), (Labeled, 8, True), (Unlabeled, 5, False)):
        ds = cls()
        if cls is Unlabeled:
            del ds.y_data
        for epoch_seed in (1, 2):
            torch.manual_seed(epoch_seed)
            ref, ref_noise = [], []
            for b in DataLoader(ds, batch_size=bs, shuffle=True, drop_last=drop):
                ref.append(b)
                ref_noise.append(torch.randn(3))               # interleaved draws, like model.encode
            torch.manual_seed(epoch_seed)
            mine, my_noise = [], []
            for b in DeviceDataLoader.from_dataset(ds, bs, shuffle=True, drop_last=drop, device="cpu"):
                mine.append(b)
                my_noise.append(torch.randn(3))
            assert len(ref) == len(mine)
            for r, m, rn, mn in zip(ref, mine, ref_noise, my_noise):
                if cls is Unlabeled:
                    assert torch.equal(r, m)
                else:
                    assert torch.equal(r[0], m[0]) and torch.equal(r[1], m[1])
                assert torch.equal(rn, mn)


def test_celeba_host_model_matches_reference_layout(golden):
    """celeba/module/model.py drop-in on the host: same-seed tensors, state_dict keys (generators unregistered), trainable
    set, arena bookkeeping, and no CPU fallback."""
    from cdgvae_b200.celeba.module.model import CDGVAE as CelebaCDGVAE
    from oracle import celeba_oracle as corc
    c = golden("celeba_linear")
    cfg = dict(c["config"], pretrained=False)
    x, y, n1, n2 = corc.synth_celeba(cfg["batch_size"], 1234, 4321)
    masks = torch.split(x[..., 3:], 1, dim=-1)
    torch.manual_seed(cfg["seed"])
    model = CelebaCDGVAE(torch.tensor(c["B"]), masks, cfg, "cpu")
    sd = model.state_dict()
    assert not any(k.startswith("decoder") for k in sd) and len(sd) == 128       # plain list, 128 registered tensors
    assert sorted(n for n, p in model.named_parameters() if p.requires_grad) == c["trainable"]
    for k, g in c["init"].items():
        t = sd[k] if k in sd else dict(model.decoder[int(k.split(".")[1])].state_dict())[k.split(".", 2)[2]]
        exact_check(t, g, k)
    assert model._n_params >= 12324 and model._frozen.numel() > 50_000_000
    w = model.decoder[0].block1.conv_1.weight_orig
    assert w.data_ptr() == model._frozen[model._foff["decoder.0.block1.conv_1.weight_orig"]:].data_ptr()   # views of the arena
    opt = torch.optim.Adam(model.parameters(), lr=1e-3)
    model.bind_optimizer(opt)
    assert model.adam_segments()[0][0] == 0
    with pytest.raises(RuntimeError):
        model(x)                                         # no CPU path


def test_last_batch_detection_counts_sized_loaders_and_reads_ahead_otherwise():
    """`xhat` is materialised for the last batch only (train.py:209).  Loaders with an exact len() are counted, so that
    batch i+1 is not requested before step i is enqueued; anything else is read one batch ahead."""
    from cdgvae_b200.modules.train import _lookahead
    from cdgvae_b200.data import DevicePrefetcher
    order = []

    class Sized:
        exact_len = True

        def __len__(self):
            return 3

        def __iter__(self):
            for i in range(3):
                order.append(("pull", i))
                yield i

    for item, last in _lookahead(Sized()):
        order.append(("step", item, last))
    assert order == [("pull", 0), ("step", 0, False), ("pull", 1), ("step", 1, False), ("pull", 2), ("step", 2, True)]
    assert list(_lookahead([7, 8])) == [(7, False), (8, True)]
    assert list(_lookahead(x for x in [7, 8, 9])) == [(7, False), (8, False), (9, True)]       # no len(): look-ahead
    assert list(_lookahead([])) == [] and list(_lookahead(x for x in [])) == []

    class LyingLen:                      # a len() nobody vouched for is not trusted
        def __len__(self):
            return 5

        def __iter__(self):
            return iter([1, 2])
    assert list(_lookahead(LyingLen())) == [(1, False), (2, True)]
    assert DevicePrefetcher([1, 2], "cuda").exact_len and not DevicePrefetcher((x for x in [1]), "cuda").exact_len


def test_tabular_arena_layout_is_the_one_the_constant_parameter_kernels_expect():
    """tabular_const.cu (loan / adult) and tab_fixed_kernel<NetCov, CP> address every weight by a compile-time offset into
    the arena and fall back to the slower shared-memory kernels when a model's layout differs.  The drop-in models must
    therefore keep producing exactly that layout (one or two 32-float slots per tensor, registration order): a silent change
    in ArenaModule would cost a factor of 1.5 - 3 without failing any parity test."""
    import synthetic_inputs as syn
    from cdgvae_b200.tabular.modules import model as M
    for name, mask in (("loan", [2, 2, 1]), ("adult", [1, 1, 3])):
        cfg = dict(dataset=name, scm="linear", flow_num=1, inverse_loop=100, lr=0.01, beta=0.01, node=3, factor=[1] * 3, input_dim=5)
        cfg["lambda"] = 10.0
        m = M.CDGVAE(syn.tabular_B(name), mask, cfg, "cpu")
        o = m._offsets
        assert m._n_params == 608                                              # TNet::NPARAMS
        assert [o["encoder.0.weight"], o["encoder.0.bias"], o["encoder.2.weight"], o["encoder.2.bias"]] == [0, 32, 64, 96]
        assert [o[f"flows.{j}.p"] for j in range(3)] == [128, 160, 192]         # TNet::flow(j)
        for k in range(3):                                                     # TNet::dw / db
            base = 224 + 128 * k
            assert [o[f"decoder.{k}.0.weight"], o[f"decoder.{k}.0.bias"], o[f"decoder.{k}.2.weight"], o[f"decoder.{k}.2.bias"]] == \
                [base, base + 32, base + 64, base + 96]
    cfg = dict(dataset="covtype", scm="linear", flow_num=1, inverse_loop=100, lr=0.01, beta=0.01, node=6, factor=[1] * 6, input_dim=8)
    cfg["lambda"] = 10.0
    m = M.CDGVAE(syn.tabular_B("covtype"), [1, 1, 2, 1, 1, 8], cfg, "cpu")
    o = m._offsets
    assert m._n_params == 1920                                                  # CovOff::NPARAMS
    assert [o[f"encoder.{i}.weight"] for i in (0, 2, 4, 6)] == [0, 64, 128, 192]  # CovOff::enc_w
    assert [o[f"encoder.{i}.bias"] for i in (0, 2, 4, 6)] == [32, 96, 160, 256]   # CovOff::enc_b
    assert [o[f"flows.{j}.p"] for j in range(6)] == [288 + 32 * j for j in range(6)]
    for k in range(6):                                                          # CovOff::dec_w / dec_b (the 7th decoder is never trained)
        for l, idx in enumerate((0, 2, 4)):
            assert o[f"decoder.{k}.{idx}.weight"] == 480 + 192 * k + 64 * l
            assert o[f"decoder.{k}.{idx}.bias"] == 480 + 192 * k + 64 * l + 32
