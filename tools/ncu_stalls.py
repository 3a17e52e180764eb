"""Read an .ncu-rep here (no GPU): headline counters of each captured launch and, for one launch, the warp-stall samples per
SASS instruction aggregated by stall reason (the `--page source` view), with the hottest instructions listed.

    python tools/ncu_stalls.py gpurun_out/x.ncu-rep [launch index] [top N]
"""
import csv
import io
import subprocess
import sys

WANT = ["gpu__time_duration.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "dram__bytes_read.sum",
        "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active"]


def I(x):
    try:
        return int(float(x))
    except Exception:
        return 0


def main():
    rep = sys.argv[1]
    launch = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        print(r[hdr.index("Kernel Name")][:70])
        for w in WANT:
            if w in hdr:
                print(f"    {w:90s} {r[hdr.index(w)]:>16s} {units[hdr.index(w)]}")
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--launch-skip", str(launch), "--launch-count", "1"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(src)))
    hdr = rows[1]
    data = [r for r in rows[2:] if len(r) == len(hdr)]
    iS, iSrc, iE = hdr.index("# Samples"), hdr.index("Source"), hdr.index("Instructions Executed")
    half = len(data) // 2
    if half and all(data[i][iSrc] == data[i + half][iSrc] for i in range(0, half, 37)):
        data = data[:half]                       # the page lists the function twice
    stall = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
    tot = sum(I(r[iS]) for r in data)
    print(f"launch {launch}: {tot} samples over {len(data)} instructions")
    for i in stall:
        s = sum(I(r[i]) for r in data)
        if s > tot * 0.01:
            print(f"    {hdr[i]:28s} {s:8d} {100 * s / tot:5.1f}%")
    for k in sorted(sorted(range(len(data)), key=lambda k: -I(data[k][iS]))[:top]):
        r = data[k]
        st = sorted(((I(r[i]), hdr[i]) for i in stall), reverse=True)[:2]
        print(f"  {k:5d} {I(r[iS]):6d} {I(r[iE]):9d}  {r[iSrc][:84]:84s} {st}")


if __name__ == "__main__":
    main()
