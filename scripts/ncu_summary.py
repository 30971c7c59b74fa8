#!/usr/bin/env python
"""Summarise an .ncu-rep (read here, no GPU): per-kernel headline metrics, stall-reason totals and opcode mix.
   python scripts/ncu_summary.py gpurun_out/prof_X.ncu-rep [kernel-regex] > profiles/X_summary.txt"""
import collections
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
pat = sys.argv[2] if len(sys.argv) > 2 else None
WANT = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'launch__registers_per_thread',
        'launch__grid_size', 'launch__block_size', 'launch__shared_mem_per_block_dynamic', 'launch__shared_mem_per_block_static',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'smsp__inst_executed.sum', 'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'sm__cycles_elapsed.max', 'lts__t_bytes.sum',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'sm__inst_executed_pipe_fma.sum', 'sm__inst_executed_pipe_alu.sum',
        'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active', 'lts__t_sector_hit_rate.pct']
raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
names = []
for r in rows[2:]:
    d = dict(zip(hdr, r))
    names.append(d['Kernel Name'])
    if pat and pat not in d['Kernel Name']:
        continue
    print('=== launch id %s: %s' % (d.get('ID'), d['Kernel Name'][:110]))
    for h, u in zip(hdr, units):
        if h in WANT:
            print('  %-62s %-14s %s' % (h, u, d[h]))
seen = set()
for nm in names:
    base = nm.split('(')[0]
    if base in seen or (pat and pat not in nm):
        continue
    seen.add(base)
    src = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--kernel-name', 'regex:' + base.split('<')[0].split('::')[-1].split()[-1]],
                         capture_output=True, text=True).stdout
    rr = list(csv.reader(io.StringIO(src)))
    # several kernels may be concatenated; take the first block
    try:
        h_i = next(i for i, r in enumerate(rr) if r and r[0] == 'Address')
    except StopIteration:
        continue
    h2 = rr[h_i]
    ix = {h: i for i, h in enumerate(h2)}
    data = []
    for r in rr[h_i + 1:]:
        if not r or r[0] in ('Kernel Name', 'Address'):
            break
        data.append(r)
    stalls = [h for h in h2 if h.startswith('stall_') and 'Not Issued' not in h]
    tot = collections.Counter()
    ops = collections.Counter()
    for r in data:
        for s_ in stalls:
            try:
                tot[s_] += int(r[ix[s_]])
            except ValueError:
                pass
        toks = r[ix['Source']].split()
        op = toks[1] if toks and toks[0].startswith('@') else (toks[0] if toks else '?')
        ops[op] += int(r[ix['Instructions Executed']] or 0)
    T = sum(tot.values()) or 1
    print('--- %s: warp-stall samples (first captured launch), %d SASS lines' % (base[:80], len(data)))
    for s_, v in tot.most_common(9):
        print('  %-26s %8d %5.1f%%' % (s_, v, 100.0 * v / T))
    te = sum(ops.values()) or 1
    print('  opcode mix (warp instructions executed, total %d):' % te)
    print('   ' + ', '.join('%s %.1f%%' % (o, 100.0 * v / te) for o, v in ops.most_common(14)))
