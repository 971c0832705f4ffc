#!/usr/bin/env python
"""Group the SASS of one kernel in an .ncu-rep into runs of equal execution count (= basic blocks / loop bodies) and
print, per run, instructions, executions, share of all executed warp instructions, stall samples and the opcode mix.
usage: python profiles/ncu_blocks.py <file.ncu-rep> <kernel-regex> [min_share_percent]"""
import collections
import csv
import io
import subprocess
import sys

rep, rx = sys.argv[1], sys.argv[2]
min_share = float(sys.argv[3]) if len(sys.argv) > 3 else 1.0
src = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--kernel-name', 'regex:' + rx], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hdr = None
inst = []
for r in rows:
    if r and r[0] == 'Address':
        if hdr is not None:
            break  # first kernel instance only
        hdr = {h: i for i, h in enumerate(r)}
        continue
    if hdr is None or len(r) < len(hdr):
        continue
    try:
        inst.append((r[hdr['Source']].strip(), int(r[hdr['Instructions Executed']]), int(r[hdr['# Samples']]),
                     float(r[hdr['Avg. Predicated-On Threads Executed']] or 0)))
    except ValueError:
        continue
total = sum(i[1] for i in inst) or 1
samples = sum(i[2] for i in inst) or 1
print('total warp instructions %d, samples %d, sass lines %d' % (total, samples, len(inst)))
blocks = []
cur = None
for k, (s, ex, sm, thr) in enumerate(inst):
    if cur is None or abs(ex - cur['ex']) > 0.02 * max(ex, cur['ex'], 1):
        cur = {'first': k, 'ex': ex, 'n': 0, 'sum': 0, 'samples': 0, 'ops': collections.Counter(), 'thr': 0.0}
        blocks.append(cur)
    cur['n'] += 1
    cur['sum'] += ex
    cur['samples'] += sm
    cur['thr'] += thr * ex
    op = s.split()[1] if s.startswith('@') else s.split()[0]
    cur['ops'][op.split('.')[0]] += 1
for b in blocks:
    share = 100.0 * b['sum'] / total
    if share < min_share:
        continue
    ops = ' '.join('%s:%d' % kv for kv in b['ops'].most_common(8))
    print('sass %4d..%4d  n=%3d  exec/inst=%9d  share=%5.1f%%  samples=%5.1f%%  active-lanes=%4.1f  %s' %
          (b['first'], b['first'] + b['n'] - 1, b['n'], b['ex'], share, 100.0 * b['samples'] / samples, b['thr'] / max(b['sum'], 1), ops))
