#!/usr/bin/env python
"""Warp-stall samples of one kernel of an .ncu-rep per CUDA source line (ncu --page source --print-source cuda,sass).
usage: ncu_lines.py REPORT.ncu-rep KERNEL_ID [TOP]"""
import csv, subprocess, sys
rep, kid = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 30
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'cuda,sass', '--kernel-id', ':::' + kid],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
fname, func, hdr, data = '', '', None, []
for r in rows:
    if not r:
        continue
    if r[0] == 'File Path':
        fname = r[1].split('/')[-1]; continue
    if r[0] == 'Function Name':
        func = r[1]; continue
    if r[0] == 'Line No':
        hdr = r; si = r.index('# Samples'); ie = r.index('Instructions Executed'); continue
    if hdr is None or len(r) <= si or r[2] != '-':
        continue
    try:
        data.append((int(r[si]), int(r[ie]), fname, r[0], r[1].strip()[:120]))
    except ValueError:
        pass
tot = sum(d[0] for d in data)
print(func[:90], '| total samples', tot, '| instructions', sum(d[1] for d in data))
for s, n, f, ln, src in sorted(data, reverse=True)[:top]:
    print('%6d %5.1f%% inst=%9d  %s:%-5s %s' % (s, 100.0 * s / max(tot, 1), n, f, ln, src))
if len(sys.argv) > 4:      # region sums: "name:lo-hi,name:lo-hi" over dense.cu-style single-file line ranges
    for spec in sys.argv[4].split(','):
        name, rng = spec.split(':'); lo, hi = map(int, rng.split('-'))
        ss = sum(d[0] for d in data if d[2] == sys.argv[5] and lo <= int(d[3]) <= hi)
        nn = sum(d[1] for d in data if d[2] == sys.argv[5] and lo <= int(d[3]) <= hi)
        print('region %-14s samples %5d (%4.1f%%) inst %9d' % (name, ss, 100.0 * ss / max(tot, 1), nn))
