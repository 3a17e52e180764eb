"""CDG-TVAE data transform (apply side).

CPU: the oracle against the reference-generated golden vectors (tests/golden/tvae_transform.json, made from the
     unmodified ClusterBasedNormalizer), host-side table construction.
GPU: the CUDA path through the C ABI against the oracle and the goldens — component indices, one-hot layout and
     category values bit-exact; normalised / recovered values bit-exact (fp64 arithmetic, same operation order);
     full-size properties (transform -> inverse round trip, one-hot structure) at 2^20 rows.
Cells whose injected uniform lies within 1e-12 of a cdf boundary (measure zero; none occur for these seeds) are the
only ones allowed to differ in the component index.
"""
import numpy as np
import pytest
import torch

from oracle import tvae_transform_oracle as orc

MARGIN = 1e-12


def golden_columns(g):
    cols = []
    for c in g["cases"]:
        cols.append(orc.continuous_column(c["attrs"], c["valid"], round_int=c["is_int"]))
    return cols


# ------------------------------------------------------------------------------------------------------------
# CPU: oracle vs the reference's own outputs
# ------------------------------------------------------------------------------------------------------------
def test_oracle_matches_reference_transform(golden):
    g = golden("tvae_transform")
    for c in g["cases"]:
        col = orc.continuous_column(c["attrs"], c["valid"], round_int=c["is_int"])
        x = np.asarray(c["raw"])
        u = np.random.RandomState(c["seed"]).random_sample(c["rows"])       # what np.random.choice consumed
        norm, comp = orc.cbn_transform(col, x, u)
        assert np.array_equal(comp, np.asarray(c["ref_component"])), c["name"]
        assert np.array_equal(norm, np.asarray(c["ref_normalized"])), c["name"]
        pp = orc.predict_proba(col, x[:8])
        assert np.abs(pp - np.asarray(c["predict_proba_head"])).max() < 1e-12


def test_oracle_matches_reference_reverse_transform(golden):
    g = golden("tvae_transform")
    for c in g["cases"]:
        col = orc.continuous_column(c["attrs"], c["valid"], round_int=c["is_int"])
        nv = int(np.sum(c["valid"]))
        data = np.zeros((c["rows"], 1 + nv), np.float32)
        data[:, 0] = np.asarray(c["inverse_in"], np.float32)
        data[np.arange(c["rows"]), 1 + np.asarray(c["ref_component"])] = 1.0
        back = orc.inverse_transform([col], data)[:, 0]
        assert np.array_equal(back, np.asarray(c["ref_inverse"])), c["name"]
        if c["is_int"]:
            assert np.array_equal(back, np.round(back))


def test_oracle_layout_and_span_table():
    cols, raw, u, z = orc.synth_table(n_cont=7, n_classes=7, rows=64, seed=3)
    out = orc.transform(cols, raw, u)
    info = orc.output_info_list(cols)
    assert out.shape[1] == sum(d for col in info for d, _ in col)
    st = 0
    for col in info:
        for dim, act in col:
            if act == "softmax":
                blk = out[:, st:st + dim]
                assert np.array_equal(blk.sum(1), np.ones(len(out), np.float32)) and set(np.unique(blk)) <= {0.0, 1.0}
            else:
                assert np.abs(out[:, st]).max() <= np.float32(0.99)
            st += dim
    back = orc.inverse_transform(cols, out)
    assert np.array_equal(back[:, -1], raw[:, -1])                          # categories survive exactly


def test_host_tables_match_oracle(golden):
    """The product's host-side table derivation (from a fitted-mixture-like object) equals the oracle's."""
    from types import SimpleNamespace
    from cdgvae_b200.tabular.modules import data_transformer as DT
    g = golden("tvae_transform")
    infos = []
    for c in g["cases"]:
        a = c["attrs"]
        n = len(a["means"])
        bgm = SimpleNamespace(weight_concentration_=(np.asarray(a["weight_concentration_a"]), np.asarray(a["weight_concentration_b"])),
                              degrees_of_freedom_=np.asarray(a["degrees_of_freedom"]), mean_precision_=np.asarray(a["mean_precision"]),
                              precisions_cholesky_=np.asarray(a["precisions_cholesky"]).reshape(n, 1, 1),
                              means_=np.asarray(a["means"]).reshape(n, 1), covariances_=np.asarray(a["covariances"]).reshape(n, 1, 1))
        gm = SimpleNamespace(_bgm_transformer=bgm, valid_component_indicator=np.asarray(c["valid"]), _dtype=np.int64 if c["is_int"] else np.float64)
        infos.append(SimpleNamespace(column_type="continuous", transform=gm))
    infos.append(SimpleNamespace(column_type="discrete", transform=SimpleNamespace(dummies=[1, 2, 3, 4, 5, 6, 7])))
    t = DT.DataTransformer.from_reference(SimpleNamespace(_column_transform_info_list=infos, dataframe=False))
    start = 0
    for i, c in enumerate(g["cases"]):
        col = orc.continuous_column(c["attrs"], c["valid"])
        d = t._cfg.col[i]
        n = d.n_all
        for name in ("mean", "std", "prec"):
            assert np.array_equal(np.asarray(getattr(d, name)[:n]), col[name]), name
        assert np.allclose(np.asarray(d.log_a[:n]), col["log_a"], rtol=1e-14, atol=0)
        assert list(d.valid_idx[:d.n_valid]) == list(np.flatnonzero(c["valid"]))
        assert d.out_start == start and d.round_int == int(c["is_int"])
        start += 1 + d.n_valid
    assert t.output_dimensions == start + 7 and t._cfg.col[len(g["cases"])].n_valid == 7
    assert [tuple(map(tuple, col)) for col in t.output_info_list] == \
        [tuple(map(tuple, col)) for col in orc.output_info_list(golden_columns(g) + [orc.discrete_column(range(1, 8))])]
    with pytest.raises(NotImplementedError):
        DT.DataTransformer().fit(np.zeros((4, 2)))
    with pytest.raises(RuntimeError):
        t.transform(np.zeros((4, 4)), device="cpu")                         # no CPU path


# ------------------------------------------------------------------------------------------------------------
# GPU parity
# ------------------------------------------------------------------------------------------------------------
def _dt(cols):
    from cdgvae_b200.tabular.modules import data_transformer as DT
    return DT.DataTransformer.from_columns(cols)


@pytest.mark.gpu
def test_gpu_transform_matches_reference_golden(golden):
    g = golden("tvae_transform")
    cols = golden_columns(g)
    rows = g["cases"][0]["rows"]
    raw = np.stack([np.asarray(c["raw"]) for c in g["cases"]], axis=1)
    u = np.stack([np.random.RandomState(c["seed"]).random_sample(rows) for c in g["cases"]])
    t = _dt(cols)
    out = t.transform(raw, uniforms=u).cpu().numpy()
    st = 0
    for i, c in enumerate(g["cases"]):
        nv = int(np.sum(c["valid"]))
        assert np.array_equal(out[:, st], np.asarray(c["ref_normalized"]).astype(np.float32)), c["name"]
        onehot = np.zeros((rows, nv), np.float32)
        onehot[np.arange(rows), np.asarray(c["ref_component"])] = 1.0
        assert np.array_equal(out[:, st + 1:st + 1 + nv], onehot), c["name"]
        st += 1 + nv
    # inverse against the reference's reverse_transform
    data = np.zeros((rows, t.output_dimensions), np.float32)
    st = 0
    for c in g["cases"]:
        nv = int(np.sum(c["valid"]))
        data[:, st] = np.asarray(c["inverse_in"], np.float32)
        data[np.arange(rows), st + 1 + np.asarray(c["ref_component"])] = 1.0
        st += 1 + nv
    back = t.inverse_transform(torch.from_numpy(data).cuda()).cpu().numpy()
    for i, c in enumerate(g["cases"]):
        assert np.array_equal(back[:, i], np.asarray(c["ref_inverse"])), c["name"]


@pytest.mark.gpu
@pytest.mark.parametrize("n_cont,n_classes,rows", [(5, 0, 1000), (7, 7, 777), (1, 0, 1), (3, 4, 129), (15, 16, 300), (2, 0, 0)])
def test_gpu_transform_matches_oracle(n_cont, n_classes, rows):
    cols, raw, u, z = orc.synth_table(n_cont, n_classes, rows, seed=11 + n_cont)
    t = _dt(cols)
    out = t.transform(raw, uniforms=u).cpu().numpy()
    ref, margin = orc.transform(cols, raw, u, return_margin=True)
    assert out.shape == ref.shape
    if rows:
        assert margin.min() > MARGIN
    assert np.array_equal(out, ref)
    # inverse, with and without the sigma draw, on a decoder-like table (soft one-hot blocks, values beyond +-1)
    rng = np.random.RandomState(5)
    data = (ref + rng.normal(0, 0.4, ref.shape)).astype(np.float32)
    sig = rng.uniform(0.01, 0.1, ref.shape[1]).astype(np.float32)
    for sigmas, normals in ((None, None), (sig, z)):
        back = t.inverse_transform(torch.from_numpy(data).cuda(), sigmas=sigmas, normals=normals).cpu().numpy()
        assert np.array_equal(back, orc.inverse_transform(cols, data, sigmas, normals))


@pytest.mark.gpu
def test_gpu_transform_fp32_filter_never_changes_a_cell(monkeypatch):
    """The fp32-filtered component draw must reproduce the all-fp64 route cell for cell (2^21 rows x 7 mixtures, far
    outliers included), and uniforms placed 1e-9 .. 1e-4 from a cdf boundary must still land on the oracle's side."""
    rows = 1 << 21
    cols, raw, u, z = orc.synth_table(7, 7, 4096, seed=21)
    t = _dt(cols)
    g = torch.Generator(device="cuda").manual_seed(4)
    idx = torch.randint(0, 4096, (rows,), device="cuda", generator=g)
    raw_d = torch.from_numpy(raw).cuda()[idx].contiguous()
    raw_d[::1013, :7] *= 50.0                                               # outliers: every component far away
    ud = torch.rand(7, rows, dtype=torch.float64, device="cuda", generator=g)
    monkeypatch.setenv("CDG_TVAE_EXACT_ONLY", "1")
    exact = t.transform(raw_d, uniforms=ud)
    monkeypatch.delenv("CDG_TVAE_EXACT_ONLY")
    fast = t.transform(raw_d, uniforms=ud)
    assert torch.equal(exact, fast)
    p_rows = 4096
    ua = u.copy()
    for cc, col in enumerate(cols[:7]):
        probs = orc.predict_proba(col, raw[:, cc])[:, col["valid"]] + 1e-6
        probs = probs / probs.sum(1, keepdims=True)
        cdf = np.cumsum(probs, axis=1)
        cdf /= cdf[:, -1:]
        b = cdf[np.arange(p_rows), np.random.RandomState(cc).randint(0, cdf.shape[1] - 1, p_rows)]
        off = 10.0 ** np.random.RandomState(100 + cc).uniform(-9, -4, p_rows) * np.random.RandomState(200 + cc).choice([-1, 1], p_rows)
        ua[cc] = np.clip(b + off, 1e-9, 1.0 - 1e-9)
    out = t.transform(raw, uniforms=ua).cpu().numpy()
    ref_a, margin_a = orc.transform(cols, raw, ua, return_margin=True)
    ok = (margin_a > MARGIN).all(axis=0)                                    # rows with every column decidable in fp64
    assert ok.mean() > 0.99
    assert np.array_equal(out[ok], ref_a[ok])


@pytest.mark.gpu
def test_gpu_transform_strided_and_unseen_category():
    cols, raw, u, z = orc.synth_table(3, 5, 500, seed=2)
    raw[::7, -1] = 42.0                                                      # value the encoder never saw: all-zero block
    t = _dt(cols)
    wide = torch.zeros(500, 9, dtype=torch.float64, device="cuda")
    wide[:, :4] = torch.from_numpy(raw).cuda()
    from cdgvae_b200 import _lib
    out = torch.empty(500, t.output_dimensions + 3, dtype=torch.float32, device="cuda").fill_(-7.0)
    ud = torch.from_numpy(u).cuda()
    _lib.check(_lib.lib().cdg_tvae_transform(t._cfg, wide.data_ptr(), 9, ud.data_ptr(), 500, out.data_ptr(), out.stride(0),
                                             torch.cuda.current_stream().cuda_stream))
    ref = orc.transform(cols, raw, u)
    assert np.array_equal(out[:, :t.output_dimensions].cpu().numpy(), ref)
    assert bool((out[:, t.output_dimensions:] == -7.0).all())                # padding columns untouched
    assert np.array_equal(ref[::7, -5:], np.zeros_like(ref[::7, -5:]))


@pytest.mark.gpu
def test_gpu_transform_invalid_config_raises():
    from cdgvae_b200 import _lib
    cols, raw, u, z = orc.synth_table(2, 0, 8, seed=1)
    t = _dt(cols)
    t._cfg.out_dim += 1
    with pytest.raises(ValueError):
        t.transform(raw, uniforms=u)


@pytest.mark.gpu
def test_gpu_gumbel_argmax_matches_reference_formula():
    from cdgvae_b200.tabular.modules.data_transformer import gumbel_argmax
    g = torch.Generator().manual_seed(3)
    logits = torch.randn(5000, 7, generator=g) * 3
    U = torch.rand(5000, 7, generator=g)
    idx, margin = orc.gumbel_argmax(logits, U)
    out = gumbel_argmax(logits.cuda(), U.cuda()).cpu().numpy().reshape(-1)
    clear = margin > 1e-4
    assert clear.mean() > 0.99
    assert np.array_equal(out[clear], idx[clear])
    wide = torch.randn(64, 49, generator=g)                                  # the covtype call site slices out[:, -7:]
    out2 = gumbel_argmax(wide[:, -7:].cuda(), U[:64].cuda()).cpu().numpy().reshape(-1)
    assert np.array_equal(out2, orc.gumbel_argmax(wide[:, -7:], U[:64])[0])


@pytest.mark.gpu
def test_gpu_transform_fullsize_properties():
    """2^20 rows of the covtype-shaped table (SURVEY §8d cfg 4): structure and round trip, no oracle needed."""
    rows = 1 << 20
    cols, raw, u, z = orc.synth_table(7, 7, 4096, seed=9)
    t = _dt(cols)
    g = torch.Generator(device="cuda").manual_seed(1)
    idx = torch.randint(0, 4096, (rows,), device="cuda", generator=g)
    raw_d = torch.from_numpy(raw).cuda()[idx].contiguous()
    ud = torch.rand(7, rows, dtype=torch.float64, device="cuda", generator=g)
    out = t.transform(raw_d, uniforms=ud)
    st = 0
    for col in t.output_info_list:
        for dim, act in col:
            blk = out[:, st:st + dim]
            if act == "softmax":
                assert bool((blk.sum(1) == 1).all()) and bool(((blk == 0) | (blk == 1)).all())
            else:
                assert float(blk.abs().max()) <= 0.99 + 1e-7
            st += dim
    # rows sharing (raw, u) with the small table agree with the oracle on it
    head = t.transform(torch.from_numpy(raw).cuda(), uniforms=u)
    assert np.array_equal(head.cpu().numpy(), orc.transform(cols, raw, u))
    # round trip: unclipped cells come back to 1e-6 relative of the component's scale (fp32 storage of the value)
    back = t.inverse_transform(out)
    assert bool((back[:, -1] == raw_d[:, -1]).all())
    unclipped = out[:, [c.out_start for c in t._cfg.col[:7]]].abs() < 0.99
    err = (back[:, :7] - raw_d[:, :7]).abs()
    assert float(err[unclipped].max()) < 1e-5
    assert float(unclipped.double().mean()) > 0.8


# ------------------------------------------------------------------------------------------------------------
# image bytes -> fp32 (modules/datasets.py:28, :42) and the prefetcher's pixel mode
# ------------------------------------------------------------------------------------------------------------
def _reference_pixels(u8):
    """modules/datasets.py:28 + :42, literally: float64 arithmetic, then torch.FloatTensor."""
    return torch.FloatTensor((np.asarray(u8).astype(float) - 127.5) / 127.5)


@pytest.mark.gpu
@pytest.mark.parametrize("n", [0, 1, 15, 256, 4097, 3 * 64 * 64 * 5])
def test_gpu_pixels_to_float_bit_exact(n):
    from cdgvae_b200 import _lib
    g = torch.Generator().manual_seed(n)
    u8 = torch.arange(256, dtype=torch.uint8) if n == 256 else torch.randint(0, 256, (n,), dtype=torch.uint8, generator=g)
    d = u8.cuda()
    out = torch.empty(n, device="cuda")
    _lib.check(_lib.lib().cdg_pixels_to_float(d.data_ptr(), n, out.data_ptr(), torch.cuda.current_stream().cuda_stream))
    assert torch.equal(out.cpu(), _reference_pixels(u8.numpy()))
    if n > 16:                                                              # unaligned views take the scalar kernel
        out2 = torch.empty(n, device="cuda")
        _lib.check(_lib.lib().cdg_pixels_to_float(d.data_ptr() + 1, n - 1, out2.data_ptr(), torch.cuda.current_stream().cuda_stream))
        assert torch.equal(out2[:n - 1].cpu(), _reference_pixels(u8.numpy()[1:]))


@pytest.mark.gpu
def test_gpu_prefetcher_pixel_mode_yields_reference_batches():
    from cdgvae_b200.data import DevicePrefetcher
    g = torch.Generator().manual_seed(0)
    batches = [(torch.randint(0, 256, (5, 8, 8, 3), dtype=torch.uint8, generator=g).pin_memory(), torch.rand(5, 5, generator=g))
               for _ in range(7)]
    seen = 0
    for (x, y), (xh, yh) in zip(DevicePrefetcher(batches, "cuda", pixels=True), batches):
        assert x.dtype == torch.float32 and x.is_cuda and torch.equal(x.cpu(), _reference_pixels(xh.numpy()))
        assert torch.equal(y.cpu(), yh)
        seen += 1
    assert seen == 7
    for (x, y), (xh, yh) in zip(DevicePrefetcher(batches, "cuda"), batches):     # default: bytes stay bytes
        assert x.dtype == torch.uint8 and torch.equal(x.cpu(), xh)


@pytest.mark.gpu
@pytest.mark.parametrize("shuffle", [True, False])
def test_gpu_device_loader_uint8_dataset_matches_the_reference_pipeline(shuffle):
    """A device-resident uint8 dataset (DeviceDataLoader(pixels=True): gather + modules/datasets.py:28 in one kernel) yields,
    bit for bit and in the same order, the batches torch's DataLoader yields over the reference's converted fp32 dataset."""
    from torch.utils.data import DataLoader, Dataset
    from cdgvae_b200.data import DeviceDataLoader
    g = torch.Generator().manual_seed(3)
    u8 = torch.randint(0, 256, (53, 8, 8, 3), dtype=torch.uint8, generator=g)
    y = torch.rand(53, 5, generator=g)

    class Ref(Dataset):                                      # modules/datasets.py:14-45 in miniature
        def __init__(self):
            self.x_data = (np.array(u8).astype(float) - 127.5) / 127.5
            self.y_data = np.array(y)

        def __len__(self):
            return len(self.x_data)

        def __getitem__(self, idx):
            return torch.FloatTensor(self.x_data[idx]), torch.FloatTensor(self.y_data[idx])

    torch.manual_seed(11)
    ref = [(a.clone(), b.clone()) for a, b in DataLoader(Ref(), batch_size=16, shuffle=shuffle)]
    torch.manual_seed(11)
    got = list(DeviceDataLoader(u8, y, batch_size=16, shuffle=shuffle, device="cuda", pixels=True))
    assert len(got) == len(ref) == 4
    for (xa, ya), (xb, yb) in zip(got, ref):
        assert xa.dtype == torch.float32 and xa.is_cuda and torch.equal(xa.cpu(), xb) and torch.equal(ya.cpu(), yb)
    with pytest.raises(TypeError):
        DeviceDataLoader(u8.float(), batch_size=4, device="cuda", pixels=True)
