#!/usr/bin/env python
"""Per-source-line share of executed warp instructions and stall samples from a .ncu-rep (needs -lineinfo + --import-source on)."""
import csv, subprocess, sys
rep = sys.argv[1]; thr = float(sys.argv[2]) if len(sys.argv) > 2 else 0.005
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
cur = None; hdr = None; agg = []
for r in rows:
    if len(r) == 2 and r[0] == 'File Path': cur = r[1].split('/')[-1]; continue
    if len(r) > 5 and r[0] == 'Line No': hdr = r; continue
    if hdr and len(r) == len(hdr) and r[0] != '':
        d = dict(zip(hdr[4:], r[4:]))
        st = {k[6:]: int(v) for k, v in d.items() if k.startswith('stall_') and not k.endswith('(Not Issued)') and v.isdigit()}
        top = max(st, key=st.get) if st else ''
        agg.append((cur, int(r[0]), r[1], int(d['Instructions Executed']), int(d['Thread Instructions Executed']), int(d['# Samples']), top, st.get(top, 0)))
tot = sum(a[3] for a in agg); tots = sum(a[5] for a in agg)
print('total warp instr', tot, 'samples', tots)
agg.sort(key=lambda a: (a[0], a[1]))
for a in agg:
    if a[3] > tot * thr or a[5] > tots * thr:
        print(f"{a[0][:14]:14s} {a[1]:4d} inst {100*a[3]/tot:5.1f}% thr/warp {a[4]/max(1,a[3]):5.1f} samp {100*a[5]/tots:5.1f}% {a[6]:>14s} {a[7]:6d} | {a[2].strip()[:80]}")
