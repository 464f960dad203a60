#!/usr/bin/env python
"""Per-CUDA-source-line share of executed warp instructions and stall samples of an `ncu --set full --import-source on`
report (needs -lineinfo):  python tools/ncu_lines.py rep.ncu-rep [min_pct] [kernel-name-regex]"""
import csv, subprocess, sys
rep = sys.argv[1]; thr = float(sys.argv[2]) if len(sys.argv) > 2 else 0.5
cmd = ['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'cuda,sass']
if len(sys.argv) > 3: cmd += ['--kernel-name', 'regex:' + sys.argv[3]]
out = subprocess.run(cmd, capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = None; lines = {}; order = []
for r in rows:
    if r and r[0] == 'Line No':
        hdr = r; ie = hdr.index('Instructions Executed'); isamp = hdr.index('# Samples')
        st = [i for i, h in enumerate(hdr) if h.startswith('stall_') and 'Not Issued' not in h]
        continue
    if hdr is None or len(r) != len(hdr) or r[0] == '': continue
    try: e = int(r[ie]); s = int(r[isamp])
    except ValueError: continue
    key = r[0]
    if key not in lines: lines[key] = [0, 0, r[1], [0] * len(st)]; order.append(key)
    L = lines[key]; L[0] += e; L[1] += s
    for k, i in enumerate(st): L[3][k] += int(r[i] or 0)
tot = sum(v[0] for v in lines.values()); ts = sum(v[1] for v in lines.values())
print('total warp instructions %d, samples %d' % (tot, ts))
for key in sorted(order, key=lambda k: int(k)):
    e, s, src, sv = lines[key]
    if 100 * e / tot >= thr or 100 * s / ts >= thr:
        top = sorted(zip(sv, [hdr[i][6:] for i in st]), reverse=True)[:2]
        print('%5s inst %5.1f%% samp %5.1f%%  %-30s %s' % (key, 100 * e / tot, 100 * s / ts, ' '.join('%s:%d' % (n, v) for v, n in top), src.strip()[:95]))
