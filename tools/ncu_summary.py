"""Summarise an .ncu-rep: key raw metrics per captured launch + stall / opcode mix of one launch."""
import csv, collections, subprocess, sys, io
rep = sys.argv[1]
which = int(sys.argv[2]) if len(sys.argv) > 2 else None
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = rows[0]
want = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'dram__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'launch__registers_per_thread', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'l1tex__t_sector_hit_rate.pct', 'lts__t_sector_hit_rate.pct', 'smsp__inst_executed.sum',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'l1tex__throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active']
idx = [(w, hdr.index(w)) for w in want if w in hdr]
for k, r in enumerate(rows[2:]):
    print(f"--- launch {k}")
    print("  " + "  ".join(f"{w.split('.')[0].replace('sm__inst_executed_pipe_','pipe_')}={float(r[i]):.1f}" for w, i in idx))
if which is not None:
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-id", f"::regex:tvl1_step:{which + 1}"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(src)))
    hi = [i for i, r in enumerate(rows[:6]) if 'Source' in r][0]
    hdr = rows[hi]; ix = {h: i for i, h in enumerate(hdr)}
    seen = set(); data = []
    for r in rows[hi + 1:]:
        if len(r) >= len(hdr) and r[0] not in seen and r[0] != hdr[0]:
            seen.add(r[0]); data.append(r)
    stalls = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
    samples = sum(int(r[ix['# Samples']] or 0) for r in data)
    tot = {s: sum(int(r[ix[s]] or 0) for r in data) for s in stalls}
    print("stalls:", {s: f"{v / samples * 100:.1f}%" for s, v in sorted(tot.items(), key=lambda kv: -kv[1])[:8]})
    ti = sum(int(r[ix['Instructions Executed']] or 0) for r in data)
    h = collections.Counter()
    for r in data:
        t = r[ix['Source']].split(); op = t[1] if t[0].startswith('@') else t[0]
        h[op.split('.')[0]] += int(r[ix['Instructions Executed']] or 0)
    print("warp instr:", ti, {op: f"{n / ti * 100:.1f}%" for op, n in h.most_common(28)})
    for r in sorted(data, key=lambda r: -int(r[ix['# Samples']] or 0))[:12]:
        print(r[ix['# Samples']].rjust(6), r[ix['Instructions Executed']].rjust(9), r[ix['Source']][:90])
