"""Summarise an ncu `--metrics gpu__time_duration.sum --csv` launch list: per-kernel totals of the last decode step."""
import collections, csv, re, sys

def load(f):
    with open(f) as fh:
        lines = [l for l in fh if not l.startswith("==")]
    rows = []
    for row in csv.DictReader(lines):
        if row.get("Metric Name") == "gpu__time_duration.sum":
            rows.append((row["Kernel Name"], float(row["Metric Value"].replace(",", "")), row.get("Grid Size", "")))
    return rows

for f in sys.argv[1:]:
    rows = load(f)
    idx = [i for i, (n, _, _) in enumerate(rows) if "embed_sum" in n]
    last = rows[idx[-1]:]
    agg = collections.defaultdict(lambda: [0, 0.0])
    for n, v, g in last:
        key = re.sub(r"\(.*", "", n)[:64]
        agg[key][0] += 1
        agg[key][1] += v
    tot = sum(v for _, v, _ in last)
    print(f"{f}: last decode step = {len(last)} kernels, {tot/1e3:.1f} us")
    for k, (c, v) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"   {v/1e3:9.1f} us  {100*v/tot:5.1f}%  x{c:4d}  avg {v/c/1e3:7.2f} us  {k}")
    # per-GEMM detail
    for n, v, g in last:
        if "gemm_tc" in n:
            print(f"      gemm grid={g} {v/1e3:.2f} us")
