#!/usr/bin/env python
"""Opcode mix per source-line range from `ncu --page source --csv --print-source cuda,sass`.
usage: ncu_opmix.py file.csv kernel_substr file_substr lo hi"""
import csv, sys, collections
path, ksub, fsub, lo, hi = sys.argv[1], sys.argv[2], sys.argv[3], int(sys.argv[4]), int(sys.argv[5])
kern = fname = hdr = None; cur_line = None; done_k = set(); first = None
mix = collections.Counter(); samples = collections.Counter()
for row in csv.reader(open(path, newline='')):
    if not row: continue
    if row[0] in ('Function Name', 'Kernel Name'):
        kern = row[1]
        if first is None and ksub in kern: first = True
        elif first and ksub in kern and fname is not None and mix: first = False  # second launch of the same kernel: stop
        continue
    if row[0] in ('File Name', 'File Path'): fname = row[1]; continue
    if row[0] == 'Line No': hdr = row; continue
    if hdr is None or kern is None or ksub not in kern or first is False: continue
    if len(row) < 10: continue
    if row[2] == '-':
        try: cur_line = int(row[0])
        except ValueError: cur_line = None
        continue
    if cur_line is None or fsub not in (fname or '') or not (lo <= cur_line <= hi): continue
    d = dict(zip(hdr, row))
    op = row[3].strip().split()[0]
    if op.startswith('@'): op = row[3].strip().split()[1]
    op = op.split('.')[0]
    try:
        mix[op] += int(d['Instructions Executed']); samples[op] += int(d['# Samples'])
    except (ValueError, KeyError): pass
tot = sum(mix.values()); ts = sum(samples.values())
print(f"total inst {tot}  samples {ts}")
for op, n in mix.most_common(25):
    print(f"{op:12s} {n:>12d} {100.0*n/tot:5.1f}%   samples {100.0*samples[op]/max(ts,1):5.1f}%")
