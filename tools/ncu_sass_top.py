#!/usr/bin/env python
"""Top SASS instructions by stall samples with their dominant stall reasons.
usage: ncu_sass_top.py file.csv kernel_substr [topN] [file_substr lo hi]"""
import csv, sys
path, ksub = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
fsub = sys.argv[4] if len(sys.argv) > 4 else None
lo = int(sys.argv[5]) if len(sys.argv) > 5 else 0
hi = int(sys.argv[6]) if len(sys.argv) > 6 else 10**9
kern = hdr = fname = cur = None; rows = {}
for row in csv.reader(open(path, newline='')):
    if not row: continue
    if row[0] in ('Function Name', 'Kernel Name'): kern = row[1]; continue
    if row[0] in ('File Name', 'File Path'): fname = row[1].split('/')[-1]; continue
    if row[0] == 'Line No': hdr = row; continue
    if hdr is None or kern is None or ksub not in kern or len(row) < 10: continue
    if row[2] == '-':
        try: cur = (fname, int(row[0]))
        except ValueError: cur = None
        continue
    if cur is None: continue
    if fsub and (fsub not in cur[0] or not (lo <= cur[1] <= hi)): continue
    d = dict(zip(hdr, row))
    try: s = int(d['# Samples'])
    except (ValueError, KeyError): continue
    rows[(row[2], cur)] = (s, row[3].strip(), d)   # keyed by address: duplicates of the same launch collapse
tot = sum(v[0] for v in rows.values())
stall_keys = [k for k in hdr if k.startswith('stall_') and 'Not Issued' not in k]
agg = {}
for (addr, cur), (s, sass, d) in rows.items():
    for k in stall_keys:
        if d[k].isdigit(): agg[k] = agg.get(k, 0) + int(d[k])
print('total samples', tot, ' stall mix:', ', '.join(f"{k[6:]} {100*v/max(tot,1):.1f}%" for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]))
for (addr, cur), (s, sass, d) in sorted(rows.items(), key=lambda kv: -kv[1][0])[:top]:
    st = {k[6:]: int(d[k]) for k in stall_keys if d[k].isdigit() and int(d[k]) > 0}
    t3 = ', '.join(f"{k} {v}" for k, v in sorted(st.items(), key=lambda kv: -kv[1])[:3])
    print(f"{100*s/max(tot,1):5.2f}% {cur[0]}:{cur[1]:<4d} {sass[:58]:58s} {t3}")
