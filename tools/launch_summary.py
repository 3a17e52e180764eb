"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel name."""
import collections, csv, re, sys
def main(path, top=25, detail=None):
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    agg = collections.defaultdict(lambda: [0, 0.0]); tot = 0.0; n = 0; rows = []
    for row in csv.DictReader(lines):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        name = re.sub(r"\(.*", "", row["Kernel Name"])
        v = float(row["Metric Value"].replace(",", "")); unit = row["Metric Unit"]
        v = v / 1e3 if unit in ("ns", "nsecond") else v * 1e3 if unit in ("ms", "msecond") else v
        agg[name][0] += 1; agg[name][1] += v; tot += v; n += 1
        rows.append((name, v, row.get("Grid Size", ""), row.get("Block Size", "")))
    print(f"{n} launches, {tot:.1f} us in total")
    for k, (c, t) in sorted(agg.items(), key=lambda x: -x[1][1])[:top]:
        print(f"{t:10.1f} us {100 * t / tot:5.1f}%  n={c:4d}  avg {t / c:8.1f} us  {k[:100]}")
    if detail:
        for name, v, g, b in rows:
            if detail in name:
                print(f"   {v:9.1f} us grid {g} {name[:80]}")
if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 25, sys.argv[3] if len(sys.argv) > 3 else None)
