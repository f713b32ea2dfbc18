"""Sum an ncu launch list (--metrics gpu__time_duration.sum --csv) per kernel: python profiles/launch_summary.py file.csv [first_kernel_substring]
With a substring, only the launches from the LAST launch group starting with that kernel are counted (the last build)."""
import collections
import csv
import sys

lines = [l for l in open(sys.argv[1]) if not l.startswith("==")]
rows = list(csv.DictReader(lines))
start = 0
if len(sys.argv) > 2:
    idx = [i for i, x in enumerate(rows) if sys.argv[2] in x["Kernel Name"]]
    groups = [i for j, i in enumerate(idx) if j == 0 or idx[j - 1] != i - 1]
    start = groups[-1] if groups else 0
agg = collections.OrderedDict()
for x in rows[start:]:
    v = float(x["Metric Value"].replace(",", ""))
    v *= {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(x["Metric Unit"], 1.0)
    a = agg.setdefault(x["Kernel Name"].split("(")[0][-70:], [0, 0.0])
    a[0] += 1
    a[1] += v
tot = sum(t for _, t in agg.values())
for n, (c, t) in agg.items():
    print(f"{t:9.3f} ms {100 * t / tot:5.1f}% x{c:4d}  {n}")
print(f"{tot:9.3f} ms total")
