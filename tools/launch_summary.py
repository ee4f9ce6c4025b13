"""Summarises an ncu `--metrics gpu__time_duration.sum --csv` launch list. Usage: python tools/launch_summary.py file.csv [--umma]"""
import collections
import csv
import re
import sys

lines = [l for l in open(sys.argv[1]) if not l.startswith("==")]
tot = collections.defaultdict(float)
cnt = collections.Counter()
seq = []
for row in csv.DictReader(lines):
    if row.get("Metric Name") != "gpu__time_duration.sum":
        continue
    v = float(row["Metric Value"].replace(",", ""))
    unit = row["Metric Unit"]
    v = v / 1e3 if unit == "ns" else (v * 1e3 if unit == "ms" else v)
    short = re.sub(r"\(.*", "", row["Kernel Name"]).replace("void ", "").replace("unnamed>::", "")
    tot[short] += v
    cnt[short] += 1
    seq.append((short, v, row["Grid Size"]))
T = sum(tot.values())
print(f"total {T:.1f} us over {len(seq)} launches")
for k, v in sorted(tot.items(), key=lambda x: -x[1]):
    print(f"{v:10.1f} us {100 * v / T:5.1f}%  n={cnt[k]:4d}  {k}")
if "--umma" in sys.argv:
    print("umma_conv launches (us, grid):")
    print([(round(v, 1), g) for s, v, g in seq if "umma" in s])
