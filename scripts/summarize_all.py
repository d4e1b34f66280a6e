"""Per-kernel totals over a whole ncu launch list (no step segmentation)."""
import collections, csv, re, sys
for f in sys.argv[1:]:
    with open(f) as fh:
        lines = [l for l in fh if not l.startswith("==")]
    agg = collections.defaultdict(lambda: [0, 0.0])
    order = []
    for row in csv.DictReader(lines):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        n = re.sub(r"\(.*", "", row["Kernel Name"])[:70]
        key = n + " grid=" + row.get("Grid Size", "") if "gemm_tc" in n else n
        agg[key][0] += 1
        agg[key][1] += float(row["Metric Value"].replace(",", ""))
    tot = sum(v for _, v in agg.values())
    print(f"{f}: {sum(c for c, _ in agg.values())} kernels, {tot/1e6:.2f} ms")
    for k, (c, v) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:40]:
        print(f"   {v/1e3:10.1f} us {100*v/tot:5.1f}%  x{c:4d}  avg {v/c/1e3:8.2f} us  {k}")
