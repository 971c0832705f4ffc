#!/usr/bin/env python
"""Summarise an .ncu-rep: per-kernel key metrics (raw page) and SASS hot spots (source page).
usage: python profiles/ncu_summary.py <file.ncu-rep> [kernel-regex]"""
import collections
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
rx = sys.argv[2] if len(sys.argv) > 2 else None
WANT = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'dram__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'smsp__inst_executed.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread', 'launch__occupancy_limit_shared_mem',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'smsp__inst_executed_op_shared_atom.sum', 'lts__t_sector_hit_rate.pct',
        'smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_membar_per_issue_active.ratio']
raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
idx = {h: i for i, h in enumerate(hdr)}
seen = set()
for r in rows[2:]:
    name = r[idx['Kernel Name']].split('(')[0]
    if name in seen or (rx and rx not in name):
        continue
    seen.add(name)
    print('==', name)
    for w in WANT:
        if w in idx:
            print('   %-86s %14s %s' % (w, r[idx[w]], units[idx[w]]))
    base = name.replace('void ', '').split('<')[0].strip()
    src = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--kernel-name', 'regex:' + base], capture_output=True, text=True).stdout
    srows = list(csv.reader(io.StringIO(src)))
    inst, cur = [], None
    for s in srows:
        if s and s[0] == 'Kernel Name':
            cur = []
            inst.append(cur)
        elif cur is not None:
            cur.append(s)
    if not inst:
        continue
    k = inst[0]
    h = {c: i for i, c in enumerate(k[0])}
    body = k[1:]
    tot = sum(int(b[h['Instructions Executed']] or 0) for b in body) or 1
    samp = sum(int(b[h['# Samples']] or 0) for b in body) or 1
    byop, bys = collections.Counter(), collections.Counter()
    for b in body:
        t = b[h['Source']].split()
        op = (t[1] if t and t[0].startswith('@') and len(t) > 1 else (t[0] if t else '?')).split('.')[0]
        byop[op] += int(b[h['Instructions Executed']] or 0)
        bys[op] += int(b[h['# Samples']] or 0)
    print('   sass lines %d, warp insts %d, samples %d' % (len(body), tot, samp))
    print('   executed by opcode:', ' '.join('%s=%.1f%%' % (o, 100.0 * c / tot) for o, c in byop.most_common(12)))
    print('   samples  by opcode:', ' '.join('%s=%.1f%%' % (o, 100.0 * c / samp) for o, c in bys.most_common(12)))
    stall_cols = [c for c in k[0] if c.startswith('stall_') and '(' not in c]
    st = collections.Counter()
    for b in body:
        for c in stall_cols:
            st[c] += int(b[h[c]] or 0)
    print('   stall samples:', ' '.join('%s=%.1f%%' % (o[6:], 100.0 * c / samp) for o, c in st.most_common(8)))
    hot = sorted(body, key=lambda b: -int(b[h['# Samples']] or 0))[:8]
    for b in hot:
        print('     hot: %-64s samp=%.1f%% exec=%.2f%%' % (b[h['Source']][:64], 100.0 * int(b[h['# Samples']] or 0) / samp,
                                                         100.0 * int(b[h['Instructions Executed']] or 0) / tot))
