"""Per-kernel totals of an `ncu --metrics gpu__time_duration.sum --csv` launch list (shares, not absolutes)."""
import collections, csv, sys

def load(path):
    with open(path) as f:
        lines = [l for l in f if l.startswith('"')]
    out = []
    for row in csv.DictReader(lines):
        t = float(row["Metric Value"].replace(",", ""))
        u = row["Metric Unit"]
        t = t / 1000 if u == "ns" else (t * 1000 if u == "ms" else t)
        out.append((row["Kernel Name"], row.get("Grid Size", ""), t))
    return out

if __name__ == "__main__":
    seq = load(sys.argv[1])
    lo = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    hi = int(sys.argv[3]) if len(sys.argv) > 3 else len(seq)
    seq = seq[lo:hi]
    per = collections.OrderedDict()
    for n, g, t in seq:
        per.setdefault((n[:90], g), []).append(t)
    tot = sum(t for _, _, t in seq)
    print(f"total {tot:.1f} us over {len(seq)} launches")
    for k, v in sorted(per.items(), key=lambda kv: -sum(kv[1])):
        print(f"{sum(v):10.1f} us {100 * sum(v) / tot:5.1f}%  n={len(v):4d}  avg={sum(v) / len(v):9.2f}  {k[0]} {k[1]}")
