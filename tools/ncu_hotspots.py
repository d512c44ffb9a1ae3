#!/usr/bin/env python
"""Summarise `ncu -i X.ncu-rep --page source --csv --print-source cuda,sass` per (file, source line): warp-stall samples,
executed instructions, shared-memory wavefronts (excess).  usage: ncu_hotspots.py file.csv [topN]"""
import csv, sys, collections
path = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
kern = None; fname = None; hdr = None
agg = collections.OrderedDict()
for row in csv.reader(open(path, newline='')):
    if not row: continue
    if row[0] == 'Function Name' or row[0] == 'Kernel Name':
        kern = row[1]; agg.setdefault(kern, {}); continue
    if row[0] in ('File Name', 'File Path'): fname = row[1].split('/')[-1]; continue
    if row[0] == 'Line No': hdr = row; continue
    if hdr is None or kern is None: continue
    d = dict(zip(hdr, row))
    if len(row) < 10 or row[2] != "-":   # sass row under a source line: skip (the source row carries the totals)
        continue
    try:
        s = int(d['# Samples']); ins = int(d['Instructions Executed'])
        shw = int(d['L1 Wavefronts Shared']); shi = int(d['L1 Wavefronts Shared Ideal'])
    except (ValueError, KeyError):
        continue
    key = (fname, int(row[0]), row[1].strip()[:90])
    a = agg[kern].setdefault(key, [0, 0, 0, 0])
    a[0] += s; a[1] += ins; a[2] += shw; a[3] += shi
seen = set()
for kern, lines in agg.items():
    tot = sum(v[0] for v in lines.values())
    if tot == 0 or (kern, tot) in seen: continue
    seen.add((kern, tot))
    print(f"===== {kern}  total samples {tot}")
    for (f, ln, src), v in sorted(lines.items(), key=lambda kv: -kv[1][0])[:top]:
        print(f"{100.0*v[0]/tot:5.1f}%  inst {v[1]:>11d}  shw {v[2]:>11d}/{v[3]:>11d}  {f}:{ln}  {src}")
# optional phase binning: PHASES env = "name:file:lo-hi,..." 
import os
ph = os.environ.get("PHASES")
if ph:
    bins = []
    for it in ph.split(","):
        nm, f, rg = it.split(":"); lo, hi = rg.split("-"); bins.append((nm, f, int(lo), int(hi)))
    for kern, lines in agg.items():
        tot = sum(v[0] for v in lines.values())
        if tot == 0: continue
        acc = collections.OrderedDict((b[0], [0, 0, 0]) for b in bins); other = [0, 0, 0]
        for (f, ln, src), v in lines.items():
            for nm, bf, lo, hi in bins:
                if f and f.startswith(bf) and lo <= ln <= hi:
                    a = acc[nm]; break
            else:
                a = other
            a[0] += v[0]; a[1] += v[1]; a[2] += v[2]
        print(f"===== phases of {kern}")
        for nm, a in list(acc.items()) + [("other", other)]:
            if a[0]: print(f"{100.0*a[0]/tot:5.1f}%  inst {a[1]:>12d} shw {a[2]:>12d}  {nm}")
