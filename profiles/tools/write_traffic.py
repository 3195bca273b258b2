"""profiles/roofline_traffic.json from an `ncu --set full` capture of the headline pruning kernel, stamped with the
sha256 of the kernel source it was taken with (bench.py quotes `roofline.traffic` only for a matching source).
usage: python profiles/tools/write_traffic.py gpurun_out/prof_prune_r2.ncu-rep"""
import csv, hashlib, io, json, os, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]
def get(name):
    i = hdr.index(name)
    v, u = float(vals[i].replace(",", "")), units[i]
    scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}[u]
    return v * scale
total = get("dram__bytes_read.sum") + get("dram__bytes_write.sum")
with open(os.path.join(ROOT, "dodonaphy_b200", "csrc", "prune.cu"), "rb") as fh:
    sha = hashlib.sha256(fh.read()).hexdigest()
out = {"workload": "c2", "kernel": vals[hdr.index("Kernel Name")], "dram_bytes_per_launch": int(total),
       "dram_bytes_read": int(get("dram__bytes_read.sum")), "dram_bytes_write": int(get("dram__bytes_write.sum")),
       "prune_cu_sha256": sha,
       "source": f"ncu --set full --clock-control none, one launch inside bench.py ({os.path.basename(rep)}; summary under profiles/r2/)"}
with open(os.path.join(ROOT, "profiles", "roofline_traffic.json"), "w") as fh:
    json.dump(out, fh, indent=1)
print(json.dumps(out))
