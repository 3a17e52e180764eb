"""Diagnostic: six free-running tabular steps against the golden, graphs on / off."""
import json, os, sys
from collections import namedtuple
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from helpers import case_setup
from cdgvae_b200.tabular.modules import model as M, train as T

for name in ("tabular_adult", "tabular_loan"):
    c = json.load(open(os.path.join(ROOT, "tests/golden", name + ".json")))
    for graphs in (True, False):
        spec, Bm, batches, cfg = case_setup(c)
        torch.manual_seed(cfg["seed"])
        model = M.CDGVAE(Bm, c["mask"], cfg, "cpu").to("cuda")
        model.use_graphs = graphs
        opt = torch.optim.Adam(model.parameters(), lr=cfg["lr"])
        DS = namedtuple("DS", ["flatten_topology"])
        for s, e in enumerate(c["steps"]):
            b = batches[s]
            model.noise_fn = lambda n, d: b["noise"]
            logs = T.train_CDGVAE(DS(c["flatten_topology"]), [(b["x"], b["y"])], model, cfg, opt, "cuda")
            dev = {k: abs(logs[k][0] - v) / (abs(v) + 1e-12) for k, v in e["logs"].items()}
            print(name, "graphs", graphs, "step", s, {k: f"{v:.2e}" for k, v in dev.items()}, flush=True)
