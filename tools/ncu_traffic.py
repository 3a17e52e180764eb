"""profiles/traffic.json for bench.py's `roofline.traffic`: dram__bytes_read.sum + dram__bytes_write.sum of ONE launch of the
dominant kernel, read (here, no GPU needed) from an `ncu --set full` report of the bench command line itself.

    python tools/ncu_traffic.py gpurun_out/x.ncu-rep <launch index in the report> <batch_per_gpu> <algorithmic bytes> "<what the launch is>"
"""
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def to_bytes(v, unit):
    v = float(v.replace(",", ""))
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}[unit]


def main():
    rep, idx, batch, alg, what = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), float(sys.argv[4]), sys.argv[5]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, r = rows[0], rows[1], rows[2 + idx]
    g = lambda k: to_bytes(r[hdr.index(k)], units[hdr.index(k)])
    out = {"batch_per_gpu": batch, "launch": what, "kernel": r[hdr.index("Kernel Name")][:120],
           "bytes": g("dram__bytes_read.sum") + g("dram__bytes_write.sum"), "algorithmic_bytes": alg,
           "duration_us_under_ncu": float(r[hdr.index("gpu__time_duration.sum")].replace(",", "")), "report": os.path.basename(rep)}
    json.dump(out, open(os.path.join(ROOT, "profiles", "traffic.json"), "w"), indent=1)
    print(out)


if __name__ == "__main__":
    main()
