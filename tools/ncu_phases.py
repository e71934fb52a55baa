#!/usr/bin/env python
"""Aggregates the per-source-line instruction and stall-sample counts of one kernel of an .ncu-rep by
line ranges of one file.  usage: ncu_phases.py rep kernel-substring file name:lo-hi [name:lo-hi ...]
Lines of other files are listed under their file name."""
import csv, subprocess, sys
rep, kern, fname = sys.argv[1], sys.argv[2], sys.argv[3]
ranges = []
for a in sys.argv[4:]:
    n, r = a.split(':'); lo, hi = r.split('-'); ranges.append((n, int(lo), int(hi)))
txt = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'cuda,sass'],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(txt.splitlines()))
f = fn = hdr = None
agg = {}
for r in rows:
    if not r: continue
    if r[0] == 'File Path': f = r[1].split('/')[-1]; continue
    if r[0] == 'Function Name': fn = r[1]; continue
    if r[0] == 'Line No': hdr = r; continue
    if hdr and len(r) == len(hdr) and kern in (fn or '') and r[2] == '-':
        d = dict(zip(hdr[4:], r[4:]))
        try:
            ie = float(d['Instructions Executed']); te = float(d['Thread Instructions Executed']); s = float(d['# Samples'])
        except ValueError:
            continue
        key = f
        if f == fname:
            key = 'other:' + fname
            for n, lo, hi in ranges:
                if lo <= int(r[0]) <= hi: key = n; break
        a = agg.setdefault(key, [0, 0, 0]); a[0] += ie; a[1] += te; a[2] += s
T = sum(v[0] for v in agg.values()); S = sum(v[2] for v in agg.values())
print(f"kernel ~{kern}: {T:.3e} warp-inst, {sum(v[1] for v in agg.values())/T:.2f} threads/inst")
for k, v in sorted(agg.items(), key=lambda x: -x[1][0]):
    print(f"{k:28s} inst {v[0]/T*100:5.1f}%  thr/inst {v[1]/max(v[0],1):5.1f}  samples {v[2]/S*100:5.1f}%")
