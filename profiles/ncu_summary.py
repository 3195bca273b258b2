"""Summarise an .ncu-rep (one kernel) into the few numbers DESIGN.md / bench.py quote.
usage: python profiles/ncu_summary.py gpurun_out/prof.ncu-rep [--top N]"""
import csv
import io
import subprocess
import sys

WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "smsp__inst_executed.sum", "sm__cycles_elapsed.avg",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "smsp__thread_inst_executed_per_inst_executed.ratio"]


def main():
    rep = sys.argv[1]
    top = int(sys.argv[sys.argv.index("--top") + 1]) if "--top" in sys.argv else 25
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    for vals in rows[2:]:
        print("== kernel:", vals[hdr.index("Kernel Name")] if "Kernel Name" in hdr else "?")
        for h, u, v in zip(hdr, units, vals):
            if h in WANT:
                print(f"  {h:75s} {v:>18s} {u}")
        st = [(float(v), h) for h, v in zip(hdr, vals) if "issue_stalled" in h and h.endswith("per_issue_active.ratio") and "not_issued" not in h]
        for v, h in sorted(st, reverse=True)[:8]:
            print(f"  stall {h.split('issue_stalled_')[1].split('_per_issue')[0]:28s} {v:8.3f} warps/issue")
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(src)))
    hdr = rows[1]
    isrc, iss, il, iex = hdr.index("Source"), hdr.index("Warp Stall Sampling (All Samples)"), hdr.index("stall_long_sb"), hdr.index("Instructions Executed")
    data = rows[2:]
    tot = sum(int(r[iss] or 0) for r in data)
    print(f"== hottest SASS instructions ({tot} samples, {len(data)} instructions)")
    for i in sorted(sorted(range(len(data)), key=lambda i: -int(data[i][iss] or 0))[:top]):
        r = data[i]
        print(f"  {i:5d} {100 * int(r[iss]) / tot:5.1f}% long_sb={r[il]:>6s} exec={r[iex]:>10s}  {r[isrc][:90]}")


if __name__ == "__main__":
    main()
