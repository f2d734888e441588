"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: share, count and mean duration per kernel.
    python tools/launch_summary.py gpurun_out/launches.csv > profiles/rNN_bench_launches_summary.txt"""
import csv, sys
from collections import defaultdict
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10]
hdr = rows[0]
ik, iv = hdr.index("Kernel Name"), hdr.index("Metric Value")
acc = defaultdict(lambda: [0, 0.0])
for r in rows[1:]:
    try:
        v = float(r[iv].replace(",", ""))
    except ValueError:
        continue
    name = r[ik].split("(")[0]
    acc[name][0] += 1
    acc[name][1] += v
tot = sum(v for _, v in acc.values())
print("ncu --metrics gpu__time_duration.sum --clock-control none -c 400 python bench.py --steps 2 --warmup 3 (BENCH_NO_EXTRAS=1): "
      "first 400 launches, device time per kernel (cold-cache, serialised: compare shares)")
for name, (n, v) in sorted(acc.items(), key=lambda kv: -kv[1][1]):
    print(f"{100 * v / tot:6.2f} %  {n:4d} launches  {v / n / 1e3:9.2f} us avg  {name}")
