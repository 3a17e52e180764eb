"""CPU oracle for the CDG-TVAE data transform (apply side).  TEST INFRASTRUCTURE ONLY.

A from-scratch numpy (float64) restatement of what the reference does on either side of train_TVAE:

    tabular/modules/numerical.py:407-445   ClusterBasedNormalizer._transform
    tabular/modules/numerical.py:447-457   ClusterBasedNormalizer._reverse_transform_helper
    tabular/modules/numerical.py:175-182   FloatFormatter._reverse_transform (rounding of integer columns)
    tabular/modules/data_transformer.py:111-129, :163-182   per-column layout, transform
    tabular/modules/data_transformer.py:131-147, :184-227   inverse_transform
    tabular/inference_tvae.py:232-235, :250-253             Gumbel-max draw of Cover_Type

Only tests/ (and tools/transform_bench.py's cpu_baseline leg) may import it; the product never does.

Third-party arithmetic on this path (absent from /root/reference, present in the image):
  * scikit-learn 1.8 `BayesianGaussianMixture.predict_proba` (numerical.py:421): restated in `bgm_tables` /
    `predict_proba` from its published algorithm (sklearn/mixture/_bayesian_mixture.py `_estimate_log_weights`
    for the dirichlet_process prior, `_estimate_log_prob`; sklearn/mixture/_base.py `_estimate_log_prob_resp`);
  * NumPy legacy `RandomState.choice(a, p=p)` (numerical.py:430-433): `cdf = p.cumsum(); cdf /= cdf[-1];
    idx = cdf.searchsorted(random_sample(), side='right')` — one uniform per call, so the reference's per-row loop
    consumes the global stream in (column, row) order; here the uniforms are an explicit argument;
  * `np.random.normal(loc, scale)` (data_transformer.py:138) = `loc + scale * standard_normal`: injected likewise;
  * rdt `OneHotEncoder` (data_transformer.py:15; the package is not installed): published behaviour restated — `dummies`
    are the distinct values in order of first appearance, transform = equality one-hot (all zeros for an unseen value),
    reverse = `dummies[argmax]`.

Parity pinning: tests/golden/make_golden_transform.py imports the UNMODIFIED reference `ClusterBasedNormalizer`
from /root/reference/tabular/modules/numerical.py, fits it on synthetic columns, runs its own `transform` /
`reverse_transform` under `np.random.seed`, checks this oracle against it cell by cell (component indices exact, values
exact) and this file's `predict_proba` against sklearn's, and commits the vectors as tests/golden/tvae_transform.json.
The data_transformer.py layer cannot be imported (rdt missing) and is pinned only through the layout its code states:
"one-hot / layout layer: parity unpinned", the normaliser beneath it: pinned.
"""
from __future__ import annotations

import numpy as np
from scipy.special import digamma

STD_MULTIPLIER = 4          # numerical.py:365
CONTINUOUS, DISCRETE = 0, 1


# --------------------------------------------------------------------------------------------------------------
# sklearn BayesianGaussianMixture (1-D, covariance_type='full', dirichlet_process) -> per-component tables
# --------------------------------------------------------------------------------------------------------------
def bgm_attrs(bgm):
    """The fitted attributes predict_proba reads, as plain float64 lists (JSON-able)."""
    a, b = bgm.weight_concentration_
    return dict(weight_concentration_a=np.asarray(a, np.float64).tolist(), weight_concentration_b=np.asarray(b, np.float64).tolist(),
                means=bgm.means_.reshape(-1).tolist(), covariances=bgm.covariances_.reshape(-1).tolist(),
                precisions_cholesky=bgm.precisions_cholesky_.reshape(-1).tolist(),
                degrees_of_freedom=bgm.degrees_of_freedom_.tolist(), mean_precision=bgm.mean_precision_.tolist(),
                weights=bgm.weights_.tolist())


def bgm_tables(attrs):
    """log(weight_k N_k(x)) = log_a[k] - 0.5 * prec[k] * (x - mean[k])^2 for one feature.

    sklearn _bayesian_mixture.py: _estimate_log_weights (dirichlet_process branch): digamma(a) - digamma(a+b) +
    cumsum of the preceding digamma(b) - digamma(a+b); _estimate_log_prob: log_gauss = -0.5 (log 2pi + dof-free
    quadratic form with precisions_cholesky) + log|chol| - 0.5 log(dof); + 0.5 (log_lambda - 1 / mean_precision) with
    log_lambda = log 2 + digamma(dof / 2); the quadratic form is scaled by dof.
    """
    a = np.asarray(attrs["weight_concentration_a"], np.float64)
    b = np.asarray(attrs["weight_concentration_b"], np.float64)
    dof = np.asarray(attrs["degrees_of_freedom"], np.float64)
    chol = np.asarray(attrs["precisions_cholesky"], np.float64)
    kappa = np.asarray(attrs["mean_precision"], np.float64)
    dg_sum = digamma(a + b)
    dg_a, dg_b = digamma(a), digamma(b)
    log_w = dg_a - dg_sum + np.hstack((0, np.cumsum(dg_b - dg_sum)[:-1]))
    # _estimate_log_gaussian_prob (n_features = 1): -0.5 (log 2pi + ((x - mu) chol)^2) + log chol
    # _estimate_log_prob: (that) - 0.5 log(dof), then + 0.5 (log_lambda - 1 / kappa); sklearn folds dof into the
    # precision it stores: precisions_cholesky_ is chol(W^-1 / dof ...) so the quadratic term needs no extra factor.
    log_lambda = np.log(2.0) + digamma(0.5 * dof)
    const = -0.5 * np.log(2.0 * np.pi) + np.log(chol) - 0.5 * np.log(dof) + 0.5 * (log_lambda - 1.0 / kappa)
    return dict(mean=np.asarray(attrs["means"], np.float64), std=np.sqrt(np.asarray(attrs["covariances"], np.float64)),
                prec=chol * chol, log_a=log_w + const)


def predict_proba(tables, x):
    """sklearn BaseMixture.predict_proba on a 1-D column x (numerical.py:421)."""
    x = np.asarray(x, np.float64).reshape(-1, 1)
    d = x - tables["mean"][None, :]
    lp = tables["log_a"][None, :] - 0.5 * (d * d * tables["prec"][None, :])
    m = lp.max(axis=1, keepdims=True)
    lse = m + np.log(np.exp(lp - m).sum(axis=1, keepdims=True))
    return np.exp(lp - lse)


# --------------------------------------------------------------------------------------------------------------
# column specs
# --------------------------------------------------------------------------------------------------------------
def continuous_column(attrs, valid, round_int=False):
    """valid: boolean mask over the fitted components (numerical.py:405: weights_ > weight_threshold)."""
    t = bgm_tables(attrs)
    return dict(kind=CONTINUOUS, valid=np.asarray(valid, bool), round_int=bool(round_int), **t)


def discrete_column(categories):
    return dict(kind=DISCRETE, categories=np.asarray(categories, np.float64))


def column_width(col):
    return 1 + int(col["valid"].sum()) if col["kind"] == CONTINUOUS else len(col["categories"])


def output_info_list(columns):
    """data_transformer.py:60-61, :78: [(1,'tanh'), (n,'softmax')] per continuous column, [(n,'softmax')] per discrete."""
    out = []
    for c in columns:
        if c["kind"] == CONTINUOUS:
            out.append([(1, "tanh"), (int(c["valid"].sum()), "softmax")])
        else:
            out.append([(len(c["categories"]), "softmax")])
    return out


# --------------------------------------------------------------------------------------------------------------
# forward
# --------------------------------------------------------------------------------------------------------------
def cbn_transform(col, x, u, return_margin=False):
    """numerical.py:407-445 for one column: (normalised value clipped to +-0.99, kept-component index).

    u: one uniform per row (the value np.random.choice would draw).
    """
    x = np.asarray(x, np.float64)
    valid = col["valid"]
    means = col["mean"].reshape(1, -1)
    stds = col["std"].reshape(1, -1)
    normalized_values = (x.reshape(-1, 1) - means) / (STD_MULTIPLIER * stds)           # :416-419
    normalized_values = normalized_values[:, valid]                                     # :420
    probs = predict_proba(col, x)[:, valid]                                             # :421-422
    sel = np.zeros(len(x), dtype=np.int64)
    margin = np.full(len(x), np.inf)
    for i in range(len(x)):                                                             # :425-433
        p = probs[i] + 1e-6
        p = p / p.sum()
        cdf = p.cumsum()
        cdf /= cdf[-1]
        sel[i] = cdf.searchsorted(u[i], side="right")
        margin[i] = np.abs(cdf - u[i]).min()
    sel = np.minimum(sel, valid.sum() - 1)
    normalized = np.clip(normalized_values[np.arange(len(x)), sel], -.99, .99)          # :435-438
    return (normalized, sel, margin) if return_margin else (normalized, sel)


def transform(columns, raw, uniforms, return_margin=False):
    """data_transformer.py:163-182 (+ :111-129): raw [B, C] float64 -> [B, D] float32.

    uniforms: [n_continuous, B] (column-major: the order of the reference's synchronous column loop, :137-145).
    """
    raw = np.asarray(raw, np.float64)
    rows = raw.shape[0]
    blocks, margins, cc = [], [], 0
    for c, col in enumerate(columns):
        w = column_width(col)
        out = np.zeros((rows, w))
        if col["kind"] == CONTINUOUS:
            v, sel, mg = cbn_transform(col, raw[:, c], uniforms[cc], return_margin=True)
            out[:, 0] = v                                                               # :119-120
            out[np.arange(rows), sel + 1] = 1.0                                         # :121-123
            margins.append(mg)
            cc += 1
        else:
            for j, cat in enumerate(col["categories"]):
                out[:, j] = (raw[:, c] == cat)
        blocks.append(out)
    res = np.concatenate(blocks, axis=1).astype(np.float32) if blocks else np.zeros((rows, 0), np.float32)
    if return_margin:
        return res, (np.stack(margins) if margins else np.zeros((0, rows)))
    return res


# --------------------------------------------------------------------------------------------------------------
# inverse
# --------------------------------------------------------------------------------------------------------------
def inverse_transform(columns, data, sigmas=None, normals=None):
    """data_transformer.py:184-227: [B, D] float32 -> raw [B, C] float64.

    sigmas: float32 [D] (model.sigma) or None; normals: [n_continuous, B] standard normals standing in for
    np.random.normal(data[:, st], sigmas[st]) (:138).
    """
    data = np.asarray(data, np.float32)
    rows = data.shape[0]
    out = np.zeros((rows, len(columns)))
    st, cc = 0, 0
    for c, col in enumerate(columns):
        w = column_width(col)
        blk = data[:, st:st + w]
        if col["kind"] == CONTINUOUS:
            comp = np.argmax(blk[:, 1:], axis=1)                                        # :135
            v = blk[:, 0].astype(np.float64)
            if sigmas is not None:
                v = v + np.float64(np.float32(sigmas[st])) * normals[cc]                # :137-139
            v = np.clip(v, -1, 1)                                                       # numerical.py:448
            std_t = col["std"][col["valid"]][comp]                                      # :453
            mean_t = col["mean"][col["valid"]][comp]                                    # :454
            r = v * STD_MULTIPLIER * std_t + mean_t                                     # :455
            if col["round_int"]:
                r = np.round(r, 0)                                                      # numerical.py:175-177
            out[:, c] = r
            cc += 1
        else:
            out[:, c] = col["categories"][np.argmax(blk, axis=1)]                       # rdt OneHotEncoder reverse
        st += w
    return out


def gumbel_argmax(logits, U):
    """tabular/inference_tvae.py:232-235, :250-253, in torch fp32 like the reference."""
    import torch
    out = torch.as_tensor(logits, dtype=torch.float32)
    U = torch.as_tensor(U, dtype=torch.float32)
    eps = 1e-20
    G = (-(U + eps).log() + eps).log()
    _, idx = (torch.nn.LogSoftmax(dim=1)(out) + G).max(dim=1, keepdim=True)
    margin = (torch.nn.LogSoftmax(dim=1)(out) + G).topk(2, dim=1).values
    return idx.reshape(-1).numpy(), (margin[:, 0] - margin[:, 1]).numpy()


# --------------------------------------------------------------------------------------------------------------
# synthetic tables (SURVEY §8d cfg 4 shapes): deterministic fitted-like mixtures without running a fit
# --------------------------------------------------------------------------------------------------------------
def synth_attrs(seed, n_all=10, n_heavy=5):
    """A plausible fitted BayesianGaussianMixture state (n_heavy real components, the rest near-empty)."""
    rng = np.random.RandomState(seed)
    n = 4000.0
    w = np.concatenate([rng.dirichlet(np.ones(n_heavy) * 3.0) * 0.999, np.full(n_all - n_heavy, 0.001 / max(1, n_all - n_heavy))])
    nk = w * n + 1e-12
    a = 1.0 + nk
    b = 0.001 + np.hstack((np.cumsum(nk[::-1])[-2::-1], 0))
    means = rng.uniform(-4, 4, n_all)
    cov = rng.uniform(0.05, 1.5, n_all)
    dof = 1.0 + nk
    kappa = 1.0 + nk
    chol = 1.0 / np.sqrt(cov)
    # sklearn's weights_ for the dirichlet process: stick-breaking expectation
    wf = a / (a + b)
    tmp = b / (a + b)
    weights = wf * np.hstack((1, np.cumprod(tmp[:-1])))
    weights /= weights.sum()
    return dict(weight_concentration_a=a.tolist(), weight_concentration_b=b.tolist(), means=means.tolist(),
                covariances=cov.tolist(), precisions_cholesky=chol.tolist(), degrees_of_freedom=dof.tolist(),
                mean_precision=kappa.tolist(), weights=weights.tolist())


def synth_table(n_cont, n_classes, rows, seed):
    """columns + raw table + injected randomness for a table of n_cont continuous columns and, when n_classes > 0,
    one trailing discrete column with values 1..n_classes (covtype's Cover_Type)."""
    rng = np.random.RandomState(seed)
    columns = []
    raw = np.zeros((rows, n_cont + (1 if n_classes else 0)))
    for c in range(n_cont):
        attrs = synth_attrs(seed * 100 + c)
        col = continuous_column(attrs, np.asarray(attrs["weights"]) > 0.005)
        comp = rng.choice(np.flatnonzero(col["valid"]), size=rows)
        raw[:, c] = col["mean"][comp] + col["std"][comp] * rng.standard_normal(rows) * 1.3
        columns.append(col)
    if n_classes:
        columns.append(discrete_column(np.arange(1, n_classes + 1)))
        raw[:, -1] = rng.randint(1, n_classes + 1, size=rows)
    uniforms = rng.random_sample((n_cont, rows))
    normals = rng.standard_normal((n_cont, rows))
    return columns, raw, uniforms, normals
