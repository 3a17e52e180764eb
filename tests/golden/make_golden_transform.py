"""Golden vectors for the CDG-TVAE data transform, made from the UNMODIFIED reference.

Run in the build container (needs /root/reference):  python tests/golden/make_golden_transform.py

Imports tabular/modules/numerical.py (ClusterBasedNormalizer) by file path, fits it on synthetic columns exactly as
DataTransformer._fit_continuous does (data_transformer.py:52-54: model_missing_values=True, max_clusters=min(len, 10),
random_state=0), runs the reference's own transform / reverse_transform under np.random.seed, cross-checks
oracle/tvae_transform_oracle.py cell by cell, and writes tests/golden/tvae_transform.json.
data_transformer.py itself cannot be imported here (it needs the rdt package, which is not installed).
"""
import importlib.util
import json
import os
import sys
import types
import warnings

import numpy as np
import pandas as pd

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import tvae_transform_oracle as orc  # noqa: E402

REF = "/root/reference/tabular/modules"


def load_reference():
    pkg = types.ModuleType("refmods")
    pkg.__path__ = [REF]
    sys.modules["refmods"] = pkg
    m = None
    for name in ["errors", "transformer_null", "transformer_base", "numerical"]:
        spec = importlib.util.spec_from_file_location("refmods." + name, f"{REF}/{name}.py")
        m = importlib.util.module_from_spec(spec)
        sys.modules["refmods." + name] = m
        spec.loader.exec_module(m)
    return m.ClusterBasedNormalizer


def synth_columns(rows):
    rng = np.random.RandomState(20231018)
    a = np.concatenate([rng.normal(-3.0, 0.5, rows * 35 // 100), rng.normal(2.0, 1.0, rows - rows * 35 // 100)])
    rng.shuffle(a)
    # integer-typed column (loan's Age / Experience style): three clusters, rounded
    b = np.concatenate([rng.normal(25, 3, rows // 3), rng.normal(45, 5, rows // 3), rng.normal(60, 2, rows - 2 * (rows // 3))])
    rng.shuffle(b)
    b = np.round(b).astype(np.int64)
    # heavy-tailed column
    c = rng.standard_t(3, rows) * 2.0 + 1.0
    return {"bimodal": a, "age": b, "heavy": c}


def main():
    warnings.simplefilter("ignore")
    CBN = load_reference()
    rows, seed = 600, 7
    cases = []
    for name, colv in synth_columns(rows).items():
        df = pd.DataFrame({name: colv})
        gm = CBN(model_missing_values=True, max_clusters=min(len(df), 10), random_state=0)   # data_transformer.py:52-53
        gm.fit(df, name)
        attrs = orc.bgm_attrs(gm._bgm_transformer)
        valid = np.asarray(gm.valid_component_indicator, bool)
        is_int = np.dtype(colv.dtype).kind == "i"
        col = orc.continuous_column(attrs, valid, round_int=is_int)

        # sklearn's predict_proba vs the restatement
        x = colv.astype(np.float64)
        pp_ref = gm._bgm_transformer.predict_proba(x.reshape(-1, 1))
        pp = orc.predict_proba(col, x)
        assert np.abs(pp - pp_ref).max() < 1e-12, np.abs(pp - pp_ref).max()

        # the reference's transform under its own RNG
        np.random.seed(seed)
        t = gm.transform(df.copy())
        ref_norm = t[f"{name}.normalized"].to_numpy()
        ref_comp = t[f"{name}.component"].to_numpy().astype(int)
        u = np.random.RandomState(seed).random_sample(rows)
        norm, comp, margin = orc.cbn_transform(col, x, u, return_margin=True)
        assert np.array_equal(comp, ref_comp), (name, np.flatnonzero(comp != ref_comp))
        assert np.array_equal(norm, ref_norm), (name, np.abs(norm - ref_norm).max())

        # the reference's reverse_transform on a perturbed transformed table
        rng = np.random.RandomState(seed + 1)
        back_in = np.stack([np.clip(ref_norm + rng.normal(0, 0.3, rows), -1.3, 1.3), ref_comp.astype(float)], axis=1)
        rdf = pd.DataFrame(back_in.copy(), columns=list(gm.get_output_sdtypes()))
        ref_back = gm.reverse_transform(rdf)[name].to_numpy()
        # oracle: the same through inverse_transform's layout (one-hot of the component)
        nv = int(valid.sum())
        data = np.zeros((rows, 1 + nv), np.float64)
        data[:, 0] = back_in[:, 0]
        data[np.arange(rows), 1 + ref_comp] = 1.0
        # inverse_transform takes the fp32 table the model emits: compare on fp32-representable inputs
        data32 = data.astype(np.float32)
        rdf32 = pd.DataFrame(np.stack([data32[:, 0].astype(np.float64), ref_comp.astype(float)], axis=1),
                             columns=list(gm.get_output_sdtypes()))
        ref_back32 = gm.reverse_transform(rdf32)[name].to_numpy()
        back = orc.inverse_transform([col], data32)[:, 0]
        assert np.array_equal(back, ref_back32.astype(np.float64)), (name, np.abs(back - ref_back32).max())
        del ref_back

        cases.append(dict(name=name, rows=rows, seed=seed, is_int=bool(is_int), attrs=attrs, valid=valid.tolist(),
                          raw=x.tolist(), ref_normalized=ref_norm.tolist(), ref_component=ref_comp.tolist(),
                          min_margin=float(margin.min()), predict_proba_head=pp_ref[:8].tolist(),
                          inverse_in=data32[:, 0].astype(np.float64).tolist(), ref_inverse=ref_back32.astype(np.float64).tolist()))
        print(name, "valid", int(valid.sum()), "min cdf margin", margin.min(), "ok")

    import sklearn
    out = dict(made_by="tests/golden/make_golden_transform.py", sklearn=sklearn.__version__, numpy=np.__version__,
               pandas=pd.__version__, cases=cases)
    with open(os.path.join(ROOT, "tests", "golden", "tvae_transform.json"), "w") as f:
        json.dump(out, f)
    print("wrote tests/golden/tvae_transform.json")


if __name__ == "__main__":
    main()
