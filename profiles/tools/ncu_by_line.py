"""Aggregate an .ncu-rep source page by CUDA source line: instructions executed and stall samples per line.
usage: python profiles/tools/ncu_by_line.py prof.ncu-rep [--top N]"""
import csv, io, subprocess, sys
rep = sys.argv[1]
top = int(sys.argv[sys.argv.index("--top") + 1]) if "--top" in sys.argv else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
agg = {}
hdr = None
cur = None
fpath = ""
for r in rows:
    if len(r) == 2 and r[0] == "File Path":
        fpath = r[1].split("/")[-1]
        continue
    if r and r[0] == "Line No":
        hdr = r
        iex, iss = hdr.index("Instructions Executed"), hdr.index("# Samples")
        continue
    if hdr is None or len(r) < len(hdr):
        continue
    if r[0] != "":
        cur = (fpath, int(r[0]), r[1].strip())
        agg.setdefault(cur, [0, 0])
    elif cur is not None:
        num = lambda t: int(t) if t.strip().isdigit() else 0
        agg[cur][0] += num(r[iex])
        agg[cur][1] += num(r[iss])
ti = sum(v[0] for v in agg.values()); ts = sum(v[1] for v in agg.values())
print(f"total instructions {ti}, samples {ts}")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
    print(f"{100*v[1]/max(ts,1):5.1f}% samples {100*v[0]/max(ti,1):5.1f}% inst  {k[0]}:{k[1]:5d}  {k[2][:100]}")
