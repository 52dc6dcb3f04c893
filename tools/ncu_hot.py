"""Summarise an ncu report of one kernel: headline metrics, stall mix, and the instructions that hold the most
stall samples (with their neighbours' opcodes).   python tools/ncu_hot.py gpurun_out/x.ncu-rep [top]"""
import csv
import subprocess
import sys

rep = sys.argv[1]
top_n = int(sys.argv[2]) if len(sys.argv) > 2 else 25
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units, vals = rows[0], rows[1], rows[2]
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "launch__registers_per_thread",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "smsp__average_warp_latency_per_inst_issued.ratio"]
for i, h in enumerate(hdr):
    if h in want or ("average_warps_issue_stalled" in h and float(vals[i].replace(",", "") or 0) > 0.2):
        print(f"{h:80s} {units[i]:10s} {vals[i]}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(src.splitlines()))
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
data = rows[2:]


def f(r, k):
    try:
        return float(r[ix[k]])
    except Exception:
        return 0.0


tot = sum(f(r, "# Samples") for r in data)
print("instructions", len(data), "samples", tot)
order = sorted(range(len(data)), key=lambda i: -f(data[i], "# Samples"))[:top_n]
for i in sorted(order):
    r = data[i]
    st = {s[6:]: f(r, s) for s in hdr if s.startswith("stall_") and "Not" not in s}
    big = sorted(st.items(), key=lambda kv: -kv[1])[:2]
    print(i, "%.2f%%" % (100 * f(r, "# Samples") / tot), r[ix["Source"]].strip()[:80], [(k, int(v)) for k, v in big])
