"""Condense `ncu --page source --csv --print-source sass` output: per-instruction executed
counts, sampled stalls and lane occupancy, plus region totals between marker opcodes."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
tot_inst = tot_samp = 0
out = []
for r in rows[2:]:
    if len(r) < len(hdr):
        continue
    ie = float(r[ix["Instructions Executed"]] or 0)
    sm = float(r[ix["# Samples"]] or 0)
    at = float(r[ix["Avg. Threads Executed"]] or 0)
    ex = float(r[ix["L1 Wavefronts Shared Excessive"]] or 0)
    tot_inst += ie; tot_samp += sm
    out.append((r[ix["Address"]], r[ix["Source"]].strip(), ie, sm, at, ex))
print("total inst %.4g samples %.4g" % (tot_inst, tot_samp))
lo = int(sys.argv[2]) if len(sys.argv) > 2 else 0
hi = int(sys.argv[3]) if len(sys.argv) > 3 else len(out)
for i, (a, s, ie, sm, at, ex) in enumerate(out[lo:hi]):
    print("%4d %s %-70s inst %6.2f%% samp %6.2f%% thr %4.1f exc %.3g" % (lo + i, a[-5:], s[:70], 100 * ie / tot_inst, 100 * sm / tot_samp, at, ex))
