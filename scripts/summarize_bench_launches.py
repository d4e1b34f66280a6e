"""Summarise an ncu `--metrics gpu__time_duration.sum --csv` launch list of `MTTS_BENCH_LAUNCHLIST=1 python bench.py`:
per-kernel launch count, total time and share of the whole hot-path step (prefill + every decode step + codec decode).

    python scripts/summarize_bench_launches.py gpurun_out/launches_bench.csv > profiles/r01_launches_bench_summary.txt
"""
import collections
import csv
import re
import sys


def load(path):
    import gzip
    with (gzip.open(path, "rt") if path.endswith(".gz") else open(path)) as fh:
        lines = [l for l in fh if not l.startswith("==")]
    for row in csv.DictReader(lines):
        if row.get("Metric Name") == "gpu__time_duration.sum":
            unit = row.get("Metric Unit", "ns")
            scale = {"ns": 1e-3, "us": 1.0, "usecond": 1.0, "nsecond": 1e-3, "ms": 1e3, "msecond": 1e3}.get(unit, 1e-3)
            yield row["Kernel Name"], float(row["Metric Value"].replace(",", "")) * scale, row.get("Grid Size", "")


def short(name):
    name = re.sub(r"\(.*", "", name)              # drop the argument list
    name = name.replace("<unnamed>::", "").replace("void ", "")
    return name[:84]


def decode_steps(rows):
    """Launches of the complete decode steps: a step starts at its embed_sum_kernel (the first one belongs to the prefill)
    and ends before the next one; the last, possibly truncated, step is dropped."""
    starts = [i for i, (n, _, _) in enumerate(rows) if "embed_sum_kernel" in n]
    if len(starts) < 3:
        return [], 0
    return rows[starts[1]:starts[-1]], len(starts) - 2


def main():
    for path in sys.argv[1:]:
        rows = list(load(path))
        dec, nsteps = decode_steps(rows)
        if nsteps:
            agg = collections.defaultdict(lambda: [0, 0.0])
            for name, us, _ in dec:
                agg[short(name)][0] += 1
                agg[short(name)][1] += us
            tot = sum(v for _, v in agg.values())
            print(f"{path}: {nsteps} complete decode steps, {len(dec) / nsteps:.0f} launches and {tot / nsteps / 1e3:.3f} ms "
                  f"per step (serialised, cold-cache)")
            for k, (c, v) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:12]:
                print(f"   {v / nsteps:10.1f} us/step  {100 * v / tot:5.1f}%  x{c / nsteps:6.1f}/step  avg {v / c:8.2f} us  {k}")
            print()
        agg = collections.defaultdict(lambda: [0, 0.0])
        total, n = 0.0, 0
        for name, us, _ in rows:
            k = short(name)
            agg[k][0] += 1
            agg[k][1] += us
            total += us
            n += 1
        print(f"{path}: {n} kernel launches, {total / 1e3:.1f} ms serialised (cold-cache ncu times)")
        ours = sum(v for k, (c, v) in agg.items() if not k.startswith("at::") and "at_cuda_detail" not in k and "cub::" not in k)
        print(f"   kernels of libmtts.so: {100 * ours / total:.1f} % of the time; torch glue (fills, copies, index ops): "
              f"{100 * (1 - ours / total):.1f} %")
        for k, (c, v) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:40]:
            print(f"   {v / 1e3:10.2f} ms  {100 * v / total:5.1f}%  x{c:6d}  avg {v / c:8.2f} us  {k}")


if __name__ == "__main__":
    main()
