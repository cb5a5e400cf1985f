"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel.

    python tools/summarize_launches.py gpurun_out/launches.csv [> profiles/rNN_summary.txt]
"""

import collections
import csv
import re
import sys

path = sys.argv[1]
with open(path) as f:
    lines = [l for l in f if not l.startswith("==")]
agg: dict[str, list[float]] = collections.defaultdict(lambda: [0, 0.0])
for row in csv.DictReader(lines):
    name = row.get("Kernel Name") or ""
    try:
        v = float((row.get("Metric Value") or "0").replace(",", ""))
    except ValueError:
        continue
    unit = row.get("Metric Unit", "")
    v *= {"usecond": 1e3, "us": 1e3, "msecond": 1e6, "ms": 1e6, "second": 1e9, "s": 1e9}.get(unit, 1.0)
    m = re.search(r"(rb200::\(anonymous namespace\)::|rb200::<unnamed>::|rb200::)(\w+)(<[^(]*)?", name)
    short = ("rb200::" + m.group(2) + (m.group(3) or "")) if m else "ATen " + re.sub(r"\(.*", "", name)[:60]
    agg[short[:80]][0] += 1
    agg[short[:80]][1] += v
total = sum(v[1] for v in agg.values())
ours = sum(v[1] for k, v in agg.items() if k.startswith("rb200::"))
print(f"# {path}: {int(sum(v[0] for v in agg.values()))} launches, {total / 1e6:.2f} ms summed kernel time "
      f"(cold-cache, serialised under ncu: compare shares, not absolutes); rb200 kernels {100 * ours / total:.1f} %")
print(f"{'ms':>10} {'share':>6} {'n':>6} {'avg us':>9}  kernel")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{v[1] / 1e6:10.3f} {100 * v[1] / total:5.1f}% {int(v[0]):6d} {v[1] / v[0] / 1e3:9.1f}  {k}")
