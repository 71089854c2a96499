#!/usr/bin/env python
"""Aggregate an `ncu --page source --csv --print-source sass,cuda` export by CUDA source line.
usage: ncu_src_hot.py file.csv <kernel substring> [top N]   -> per (file, line): share of samples, share of instructions, top stalls"""
import csv
import os
import sys
path, want = sys.argv[1], (sys.argv[2] if len(sys.argv) > 2 else '')
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
rows = list(csv.reader(open(path)))
agg = []
i = 0
while i < len(rows):
    if rows[i] and rows[i][0] == 'File Path':
        fpath, fn, hdr = rows[i][1], rows[i + 1][1], rows[i + 2]
        j = i + 3
        samp, inst = hdr.index('# Samples'), hdr.index('Instructions Executed')
        stall_cols = [k for k, h in enumerate(hdr) if h.startswith('stall_') and 'Not Issued' not in h]
        while j < len(rows) and not (rows[j] and rows[j][0] == 'File Path'):
            r = rows[j]
            j += 1
            if want not in fn or not r or not r[0].strip().isdigit():
                continue
            try:
                s, n = float(r[samp] or 0), float(r[inst] or 0)
            except ValueError:
                continue
            st = sorted(((float(r[k] or 0), hdr[k][6:]) for k in stall_cols), reverse=True)[:3]
            agg.append((s, n, os.path.basename(fpath), int(r[0]), r[1].strip()[:100], st))
        i = j
    else:
        i += 1
tot = sum(a[0] for a in agg) or 1
totn = sum(a[1] for a in agg) or 1
print(f'kernel ~{want}: samples {tot:.0f}, warp instructions {totn:.0f}')
byfile = {}
for a in agg:
    f = byfile.setdefault(a[2], [0, 0]); f[0] += a[0]; f[1] += a[1]
for f, (s, n) in byfile.items():
    print(f'  file {f:24s} {s / tot * 100:5.1f}% samples {n / totn * 100:5.1f}% instr')
for s, n, f, ln, src, st in sorted(agg, reverse=True)[:top]:
    sts = ' '.join(f'{h}:{v / max(s, 1) * 100:.0f}%' for v, h in st)
    print(f'{s / tot * 100:5.1f}%s {n / totn * 100:5.1f}%i {f[:14]:14s}:{ln:<4d} {src[:90]:90s} [{sts}]')
