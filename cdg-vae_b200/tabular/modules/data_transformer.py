"""Apply side of the reference's tabular/modules/data_transformer.py on the B200.

    DataTransformer.transform(raw)                   data_transformer.py:163-182  -> cdg_tvae_transform
    DataTransformer.inverse_transform(data, sigmas)  data_transformer.py:184-227  -> cdg_tvae_inverse_transform
    gumbel_argmax(logits)                            tabular/inference_tvae.py:232-235, :250-253 -> cdg_gumbel_argmax

Same attribute names as the reference (`output_info_list`, `output_dimensions`, `SpanInfo`), so `train_TVAE(
transformer.output_info_list, ...)` and `config["input_dim"] = transformer.output_dimensions` (main_tvae.py:131, :205)
read the same things.  FITTING is not done here (BayesianGaussianMixture / category discovery are host-side data
preparation, SURVEY §8f row 4 scopes the apply side): build the object from a transformer the reference has fitted
(`from_reference`) or from explicit per-column tables (`from_columns`).

Tables live in device memory as torch tensors: raw tables float64 `[rows, n_columns]`, transformed tables float32
`[rows, output_dimensions]`.  The reference draws its randomness from NumPy's global generator on the host
(numerical.py:430, data_transformer.py:138); here it is an argument (`uniforms`, `normals`: `[n_continuous, rows]`
float64, the order of the reference's column loop) and is drawn on the device when omitted.
"""
from collections import namedtuple

from ... import _lib

SpanInfo = namedtuple('SpanInfo', ['dim', 'activation_fn'])          # data_transformer.py:18
CONTINUOUS, DISCRETE = 0, 1


def _bgm_tables(bgm):
    """Per-component tables of a fitted sklearn BayesianGaussianMixture (1 feature, dirichlet_process prior):
    log(weight_k N_k(x)) = log_a[k] - 0.5 prec[k] (x - mean[k])^2 — what predict_proba (numerical.py:421) evaluates,
    with the data-independent digamma terms of _estimate_log_weights / _estimate_log_prob folded into log_a."""
    import numpy as np
    from scipy.special import digamma
    a, b = (np.asarray(v, np.float64) for v in bgm.weight_concentration_)
    dof = np.asarray(bgm.degrees_of_freedom_, np.float64)
    chol = np.asarray(bgm.precisions_cholesky_, np.float64).reshape(-1)
    kappa = np.asarray(bgm.mean_precision_, np.float64)
    dsum = digamma(a + b)
    log_w = digamma(a) - dsum + np.hstack((0, np.cumsum(digamma(b) - dsum)[:-1]))
    log_lambda = np.log(2.0) + digamma(0.5 * dof)
    log_a = log_w - 0.5 * np.log(2.0 * np.pi) + np.log(chol) - 0.5 * np.log(dof) + 0.5 * (log_lambda - 1.0 / kappa)
    return dict(mean=np.asarray(bgm.means_, np.float64).reshape(-1), std=np.sqrt(np.asarray(bgm.covariances_, np.float64).reshape(-1)),
                prec=chol * chol, log_a=log_a)


class DataTransformer(object):
    """Mode-specific normalisation of continuous columns, one-hot of discrete columns (apply only)."""

    def __init__(self, max_clusters=10, weight_threshold=0.005):
        self._max_clusters = max_clusters
        self._weight_threshold = weight_threshold
        self._columns = None

    # ---- construction ------------------------------------------------------------------------------------
    def fit(self, raw_data, discrete_columns=(), random_state=0):
        raise NotImplementedError(
            "fitting (BayesianGaussianMixture, category discovery) is host-side data preparation and is not part of this "
            "library: fit the reference's DataTransformer and pass it to DataTransformer.from_reference(), or use "
            "DataTransformer.from_columns()")

    @classmethod
    def from_reference(cls, fitted):
        """`fitted`: a reference DataTransformer after .fit() (reads _column_transform_info_list, data_transformer.py:109)."""
        import numpy as np
        cols = []
        for info in fitted._column_transform_info_list:
            if info.column_type == 'continuous':
                gm = info.transform
                t = _bgm_tables(gm._bgm_transformer)
                cols.append(dict(kind=CONTINUOUS, valid=np.asarray(gm.valid_component_indicator, bool),
                                 round_int=np.dtype(getattr(gm, "_dtype", float)).kind == 'i', **t))
            else:
                cols.append(dict(kind=DISCRETE, categories=np.asarray(info.transform.dummies, np.float64)))
        self = cls.from_columns(cols)
        self.dataframe = getattr(fitted, "dataframe", False)
        return self

    @classmethod
    def from_columns(cls, columns):
        """columns: list of dicts — continuous: kind=0, mean/std/prec/log_a (all fitted components), valid (bool mask),
        round_int; discrete: kind=1, categories (float values in one-hot order)."""
        import numpy as np
        self = cls()
        cfg = _lib.TvaeTransformConfig()
        if not 1 <= len(columns) <= _lib.MAX_TCOL:
            raise ValueError(f"a table has 1..{_lib.MAX_TCOL} columns, got {len(columns)}")
        start, info = 0, []
        for c, col in enumerate(columns):
            d = cfg.col[c]
            d.kind, d.out_start = int(col["kind"]), start
            if col["kind"] == CONTINUOUS:
                valid = np.flatnonzero(np.asarray(col["valid"], bool))
                n_all = len(col["mean"])
                if n_all > _lib.MAX_TCOMP or len(valid) < 1:
                    raise ValueError(f"column {c}: {len(valid)} of {n_all} components (at most {_lib.MAX_TCOMP}, at least 1 kept)")
                d.n_all, d.n_valid, d.round_int = n_all, len(valid), int(bool(col.get("round_int", False)))
                for j, k in enumerate(valid):
                    d.valid_idx[j] = int(k)
                for k in range(n_all):
                    d.mean[k], d.std[k], d.prec[k], d.log_a[k] = (float(col[n][k]) for n in ("mean", "std", "prec", "log_a"))
                info.append([SpanInfo(1, 'tanh'), SpanInfo(len(valid), 'softmax')])           # data_transformer.py:60
                start += 1 + len(valid)
            elif col["kind"] == DISCRETE:
                cats = np.asarray(col["categories"], np.float64)
                if not 1 <= len(cats) <= _lib.MAX_TCAT:
                    raise ValueError(f"column {c}: {len(cats)} categories (1..{_lib.MAX_TCAT})")
                d.n_valid = len(cats)
                for j, v in enumerate(cats):
                    d.category[j] = float(v)
                info.append([SpanInfo(len(cats), 'softmax')])                                  # data_transformer.py:78
                start += len(cats)
            else:
                raise ValueError(f"column {c}: unknown kind {col['kind']!r}")
        cfg.n_col, cfg.out_dim = len(columns), start
        self._cfg = cfg
        self._columns = columns
        self._n_cont = sum(1 for col in columns if col["kind"] == CONTINUOUS)
        self.output_info_list = info
        self.output_dimensions = start
        self.dataframe = False
        return self

    # ---- apply -------------------------------------------------------------------------------------------
    def _raw_tensor(self, raw_data, device):
        import torch
        if hasattr(raw_data, "to_numpy"):
            raw_data = raw_data.to_numpy()
        t = torch.as_tensor(raw_data)
        return t.to(device=device, dtype=torch.float64).contiguous()

    def transform(self, raw_data, uniforms=None, device="cuda"):
        """raw_data [rows, n_columns] (array / DataFrame / tensor) -> float32 tensor [rows, output_dimensions] on `device`."""
        import torch
        dev = _lib.require_cuda(device)
        raw = self._raw_tensor(raw_data, dev)
        rows = raw.shape[0]
        if raw.dim() != 2 or raw.shape[1] != self._cfg.n_col:
            raise ValueError(f"raw table must be [rows, {self._cfg.n_col}], got {tuple(raw.shape)}")
        if uniforms is None:
            uniforms = torch.rand(self._n_cont, rows, dtype=torch.float64, device=dev)
        u = torch.as_tensor(uniforms).to(device=dev, dtype=torch.float64).contiguous()
        if tuple(u.shape) != (self._n_cont, rows):
            raise ValueError(f"uniforms must be [{self._n_cont}, {rows}], got {tuple(u.shape)}")
        out = torch.empty(rows, self.output_dimensions, dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            _lib.check(_lib.lib().cdg_tvae_transform(self._cfg, raw.data_ptr(), raw.stride(0) if rows else self._cfg.n_col,
                                                     u.data_ptr(), rows, out.data_ptr(), self.output_dimensions,
                                                     torch.cuda.current_stream().cuda_stream))
        return out

    def inverse_transform(self, data, sigmas=None, normals=None):
        """data float32 [rows, output_dimensions] on the GPU -> float64 tensor [rows, n_columns] (same device).

        With `sigmas` (model.sigma) every continuous value is first redrawn as N(value, sigmas[start]) (:136-140);
        `normals` [n_continuous, rows] injects the standard normals of that draw.
        """
        import torch
        if not torch.is_tensor(data):
            data = torch.as_tensor(data)
        dev = _lib.require_cuda(data.device if data.is_cuda else "cuda")
        data = data.detach().to(device=dev, dtype=torch.float32).contiguous()
        rows = data.shape[0]
        if data.dim() != 2 or data.shape[1] != self.output_dimensions:
            raise ValueError(f"data must be [rows, {self.output_dimensions}], got {tuple(data.shape)}")
        sg = nz = None
        if sigmas is not None:
            sg = torch.as_tensor(sigmas).detach().to(device=dev, dtype=torch.float32).contiguous()
            if sg.numel() != self.output_dimensions:
                raise ValueError(f"sigmas must have {self.output_dimensions} entries")
            if normals is None:
                normals = torch.randn(self._n_cont, rows, dtype=torch.float64, device=dev)
            nz = torch.as_tensor(normals).to(device=dev, dtype=torch.float64).contiguous()
            if tuple(nz.shape) != (self._n_cont, rows):
                raise ValueError(f"normals must be [{self._n_cont}, {rows}], got {tuple(nz.shape)}")
        out = torch.empty(rows, self._cfg.n_col, dtype=torch.float64, device=dev)
        with torch.cuda.device(dev):
            _lib.check(_lib.lib().cdg_tvae_inverse_transform(
                self._cfg, data.data_ptr(), self.output_dimensions, None if sg is None else sg.data_ptr(),
                None if nz is None else nz.data_ptr(), rows, out.data_ptr(), self._cfg.n_col,
                torch.cuda.current_stream().cuda_stream))
        return out


def gumbel_argmax(logits, uniforms=None):
    """Class draw of tabular/inference_tvae.py:232-235, :250-253: argmax(log_softmax(logits) + log(-log(U + eps) + eps)).
    logits float32 [rows, n_class] on the GPU -> int64 [rows, 1]."""
    import torch
    dev = _lib.require_cuda(logits.device if torch.is_tensor(logits) and logits.is_cuda else "cuda")
    x = torch.as_tensor(logits).detach().to(device=dev, dtype=torch.float32)
    if x.dim() != 2:
        raise ValueError("logits must be [rows, n_class]")
    if x.stride(1) != 1:
        x = x.contiguous()
    rows, n = x.shape
    if uniforms is None:
        uniforms = torch.rand(rows, n, device=dev)
    u = torch.as_tensor(uniforms).to(device=dev, dtype=torch.float32).contiguous()
    if tuple(u.shape) != (rows, n):
        raise ValueError(f"uniforms must be [{rows}, {n}]")
    out = torch.empty(rows, 1, dtype=torch.int64, device=dev)
    with torch.cuda.device(dev):
        _lib.check(_lib.lib().cdg_gumbel_argmax(x.data_ptr(), x.stride(0) if rows else n, n, u.data_ptr(), rows, out.data_ptr(),
                                                torch.cuda.current_stream().cuda_stream))
    return out
