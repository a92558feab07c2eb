"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel count, total time and share.
Usage: python scripts/summarize_launches.py gpurun_out/launches.csv [skip_fraction] > profiles/xxx.md"""
import csv, re, sys
from collections import defaultdict

path = sys.argv[1]
rows = []
with open(path, newline="") as f:
    lines = [l for l in f if not l.startswith("==")]
reader = csv.DictReader(lines)
for r in reader:
    if r.get("Metric Name") != "gpu__time_duration.sum":
        continue
    val = float(r["Metric Value"].replace(",", ""))
    unit = r.get("Metric Unit", "ns")
    ns = val * {"ns": 1, "us": 1e3, "ms": 1e6, "s": 1e9}.get(unit, 1)
    rows.append((int(r["ID"]), r["Kernel Name"], ns))
rows.sort()
# keep the last timed step only: launches after the last bl adam kernel of the warm-up
adam = [i for i, (_, n, _) in enumerate(rows) if "adam_kernel" in n]
start = adam[-2] + 1 if len(adam) >= 2 else 0
end = adam[-1] + 1 if adam else len(rows)
step = rows[start:end]
agg = defaultdict(lambda: [0, 0.0])
for _, name, ns in step:
    short = re.sub(r"\(.*", "", name)
    short = re.sub(r"^void ", "", short)
    short = short[:110]
    agg[short][0] += 1
    agg[short][1] += ns
total = sum(v[1] for v in agg.values())
print(f"# launch list summary: {path}\n")
print(f"launches in the profiled train step: {len(step)}; sum of kernel durations {total/1e6:.1f} ms "
      f"(serialised, cold-cache under ncu: compare SHARES, not absolutes)\n")
print("| kernel | launches | total ms | share |\n|---|---:|---:|---:|")
for name, (cnt, ns) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:40]:
    print(f"| `{name}` | {cnt} | {ns/1e6:.2f} | {100*ns/total:.1f}% |")
