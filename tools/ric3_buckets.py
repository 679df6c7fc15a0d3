#!/usr/bin/env python
"""Bucket tools/ncu_lines.py output of k_riccati3 into the phases of riccati3.cuh (reads the marker comments)."""
import re, sys
src = open('simple-mpc_b200/csrc/riccati3.cuh').read().splitlines()
marks = []
for n, l in enumerate(src, 1):
    if 'SMPC_DEV void chol_reg' in l: marks.append((n - 16, 'chol_reg'))
    elif 'SMPC_DEV void axpy_row' in l: marks.append((n - 4, 'sweeps'))
    elif 'SMPC_DEV void riccati3_backward_inst' in l: marks.append((n - 1, 'prologue/terminal'))
    elif re.search(r'// ---- (0|1|2|3|4|5)\.', l): marks.append((n, l.strip()[8:40]))
    elif '// ---- gains' in l: marks.append((n, 'gains out'))
    elif 'Sigma = mu I + W^T W (lower) and G' in l: marks.append((n, 'Sigma,G build'))
    elif '// Y <- Y + W G' in l: marks.append((n, 'Y += W G'))
    elif '// symmetrise P' in l: marks.append((n, 'symmetrise'))
    elif 'for (int e = tid; e < NU * NRY; e += SMPC_NT)' in l: marks.append((n - 1, 'YW init'))
marks.sort()
acc = {}
wfa = {}
wfia = {}
for l in sys.stdin:
    m = re.match(r'\s*([\d.]+)%\s+(\d+) samples\s+(\d+) inst\s+(\S+):(\d+)', l)
    if not m: continue
    mw = re.search(r'smem_wavefronts (\d+) ideal (\d+)', l)
    pct, f, ln = float(m.group(1)), m.group(4), int(m.group(5))
    key = 'other:' + f
    if f == 'riccati3.cuh':
        key = 'header'
        for n, name in marks:
            if ln >= n: key = name
    acc[key] = acc.get(key, 0) + pct
    if mw:
        wfa[key] = wfa.get(key, 0) + int(mw.group(1))
        wfia[key] = wfia.get(key, 0) + int(mw.group(2))
tw = max(sum(wfa.values()), 1)
for k, v in sorted(acc.items(), key=lambda x: -x[1]): print('%6.2f%%  smem wavefronts %5.1f%% (ideal/actual %.2f)  %s' % (v, 100.0 * wfa.get(k, 0) / tw, wfia.get(k, 0) / max(wfa.get(k, 0), 1), k))
