"""Key metrics + SASS region summary of an .ncu-rep (first kernel).  usage: ncu_summary.py rep [chunk]"""
import csv, subprocess, sys
rep = sys.argv[1]
chunk = int(sys.argv[2]) if len(sys.argv) > 2 else 48
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units, vals = rows[0], rows[1], rows[2]
keys = ['gpu__time_duration.sum', 'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'smsp__thread_inst_executed_per_inst_executed.ratio', 'launch__registers_per_thread', 'launch__occupancy_limit_registers',
        'launch__occupancy_limit_shared_mem', 'smsp__inst_executed.sum', 'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'smsp__cycles_active.avg',
        'launch__grid_size', 'launch__block_size', 'launch__shared_mem_per_block_dynamic', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed']
for i, h in enumerate(hdr):
    if h in keys or ('issue_stalled' in h and 'per_issue_active' in h and float(vals[i] or 0) > 0.1):
        print(f"{h},{units[i]},{vals[i]}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(src.splitlines()))
hdr = rows[1]; ix = {h: i for i, h in enumerate(hdr)}
body = []
for r in rows[2:]:
    if len(r) < len(hdr):
        if body: break          # next kernel of the report
        continue
    if r[ix['Source']] == 'Source':
        break
    body.append(r)
data = [(r[ix['Source']].strip(), float(r[ix['Instructions Executed']] or 0), float(r[ix['# Samples']] or 0),
         float(r[ix['Avg. Threads Executed']] or 0)) for r in body]
T = sum(x[1] for x in data); S = sum(x[2] for x in data)
print("# SASS regions (instructions executed %, stall samples %, avg active threads)")
for i in range(0, len(data), chunk):
    blk = data[i:i + chunk]
    ie = sum(b[1] for b in blk); sm = sum(b[2] for b in blk)
    if ie / T < 0.002 and sm / S < 0.002:
        continue
    thr = sum(b[1] * b[3] for b in blk) / max(ie, 1)
    ops = {}
    for b in blk:
        t = b[0].split()
        if not t: continue
        op = (t[1] if t[0].startswith('@') else t[0]).split('.')[0]
        ops[op] = ops.get(op, 0) + b[1]
    top = ' '.join(f"{k}:{100*v/T:.1f}" for k, v in sorted(ops.items(), key=lambda kv: -kv[1])[:7])
    print(f"{i:5d} inst {100*ie/T:5.1f}% samp {100*sm/S:5.1f}% thr {thr:4.1f}  {top}")
