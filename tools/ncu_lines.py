#!/usr/bin/env python
"""Per-source-line instruction/stall breakdown of one kernel of an .ncu-rep (needs -lineinfo and
--import-source on).  usage: tools_ncu_lines.py rep kernel-substring [top]"""
import csv, subprocess, sys, collections
rep, kern = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
txt = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'cuda,sass'],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(txt.splitlines()))
f = fn = None
hdr = None
agg = {}
for r in rows:
    if not r: continue
    if r[0] == 'File Path': f = r[1].split('/')[-1]; continue
    if r[0] == 'Function Name': fn = r[1]; continue
    if r[0] == 'Line No': hdr = r; continue
    if hdr and len(r) == len(hdr) and kern in (fn or '') and r[2] == '-':   # source-line aggregate row
        d = dict(zip(hdr[4:], r[4:]))
        try:
            ie = float(d['Instructions Executed']); te = float(d['Thread Instructions Executed']); s = float(d['# Samples'])
        except ValueError:
            continue
        if ie > 0 or s > 0:
            k = (f, int(r[0]))
            a = agg.setdefault(k, [0, 0, 0, r[1].strip()[:100]])
            a[0] += ie; a[1] += te; a[2] += s
T = sum(v[0] for v in agg.values()); S = sum(v[2] for v in agg.values())
print(f"kernel ~{kern}: {T:.3e} warp-inst, {sum(v[1] for v in agg.values())/T:.2f} threads/inst, {S:.0f} samples")
for k, v in sorted(agg.items(), key=lambda x: -x[1][2])[:top]:
    print(f"{k[0]:18s} L{k[1]:<4d} inst {v[0]/T*100:5.1f}%  thr/inst {v[1]/max(v[0],1):5.1f}  samples {v[2]/S*100:5.1f}% | {v[3]}")
