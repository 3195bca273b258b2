"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel.
usage: python profiles/launch_summary.py gpurun_out/launches.csv"""
import collections
import csv
import sys

lines = [l for l in open(sys.argv[1]) if not l.startswith("==")]
agg = collections.defaultdict(list)
for row in csv.DictReader(lines):
    if row.get("Metric Name") == "gpu__time_duration.sum":
        v = float(row["Metric Value"])
        unit = row["Metric Unit"]
        us = v / 1e3 if unit.startswith("n") else (v * 1e3 if unit.startswith("m") else v)
        agg[row["Kernel Name"][:70]].append(us)
tot = sum(sum(v) for v in agg.values())
print(f"{'kernel':72s} {'n':>4s} {'mean us':>10s} {'share':>7s}")
for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
    print(f"{k:72s} {len(v):4d} {sum(v) / len(v):10.1f} {sum(v) / tot:7.4f}")
