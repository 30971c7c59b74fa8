#!/usr/bin/env python
"""Hot spots of one kernel in an .ncu-rep (read here, no GPU): SASS lines with the most warp-stall samples, with the running share,
plus the dominant stall reason per line.   python scripts/ncu_hot.py rep.ncu-rep [top]"""
import csv, io, subprocess, sys
rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 60
src = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv'], capture_output=True, text=True).stdout
rr = list(csv.reader(io.StringIO(src)))
h_i = next(i for i, r in enumerate(rr) if r and r[0] == 'Address')
hdr = rr[h_i]
ix = {h: i for i, h in enumerate(hdr)}
stalls = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
data = [r for r in rr[h_i + 1:] if r and len(r) == len(hdr)]
tot = sum(int(r[ix['# Samples']] or 0) for r in data)
print("kernel:", rr[0][1][:100]); print("total samples", tot, "SASS lines", len(data))
order = sorted(range(len(data)), key=lambda i: -int(data[i][ix['# Samples']] or 0))[:top]
for i in sorted(order):
    r = data[i]
    n = int(r[ix['# Samples']] or 0)
    best = max(stalls, key=lambda s_: int(r[ix[s_]] or 0))
    print("%5d %5.1f%%  exec %8s  %-22s %s" % (i, 100.0 * n / tot, r[ix['Instructions Executed']], best, r[ix['Source']].strip()[:90]))
print("--- samples per 100-line window")
for a in range(0, len(data), 100):
    n = sum(int(r[ix['# Samples']] or 0) for r in data[a:a + 100])
    if n * 200 > tot:
        print("lines %5d-%5d: %5.1f%%" % (a, a + 99, 100.0 * n / tot))
