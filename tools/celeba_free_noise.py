"""How far do two runs of the UNMODIFIED CelebA reference drift apart over six free-running steps when only the BLAS
thread count changes?  (Build container only: imports /root/reference through tests/golden/make_golden_celeba.py.)
The answer is the noise floor tests/test_free_running_gpu.py's CelebA bound is set against; output committed as
profiles/r02_celeba_free_running_noise.json."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
import make_golden_celeba as mg  # noqa: E402


def run(threads, steps=6, batch=2):
    torch.set_num_threads(threads)
    c = mg.celeba_free_case(f"t{threads}", batch, steps)
    return [s["logs"] for s in c["steps"]], c["final_params"]


def main():
    out = {}
    base_logs, base_par = run(os.cpu_count())
    for t in (1, 2):
        logs, par = run(t)
        dev = [max(abs(a[k] - b[k]) / (abs(b[k]) + 1e-12) for k in a if abs(b[k]) > 1e-6) for a, b in zip(logs, base_logs)]
        pdev = max(max(abs(x - y) for x, y in zip(par[n]["val"], base_par[n]["val"])) / (base_par[n]["absmax"] + 1e-30) for n in par)
        out[f"threads_{t}_vs_{os.cpu_count()}"] = {"log_rel_dev_per_step": dev, "param_sample_dev_max": pdev}
        print(t, dev, pdev, flush=True)
    json.dump(out, open(os.path.join(ROOT, "profiles", "r02_celeba_free_running_noise.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
