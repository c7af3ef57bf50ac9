#!/usr/bin/env python3
"""Summarise an .ncu-rep: per-kernel key metrics + top stall-sampled SASS lines.  usage: ncu_top.py rep [kernel-regex] [n]"""
import csv, io, subprocess, sys
rep = sys.argv[1]; kre = sys.argv[2] if len(sys.argv) > 2 else None; n = int(sys.argv[3]) if len(sys.argv) > 3 else 25
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
want = ['gpu__time_duration.sum', 'sm__cycles_elapsed.max', 'smsp__inst_executed.sum', 'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active', 'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active',
        'dram__bytes_read.sum', 'dram__bytes_write.sum', 'lts__t_bytes.sum', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'launch__registers_per_thread', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'lts__t_sectors_op_red.sum', 'lts__t_sectors_op_atom.sum']
for r in rows[2:]:
    print('====', r[hdr.index('Kernel Name')][:60])
    for w in want:
        if w in hdr: print(f'  {w} = {r[hdr.index(w)]} {units[hdr.index(w)]}')
    st = []
    for i, h in enumerate(hdr):
        if 'issue_stalled' in h and h.endswith('per_issue_active.ratio'):
            try: st.append((float(r[i]), h.split('issue_stalled_')[1].replace('_per_issue_active.ratio', '')))
            except: pass
    print('  stalls/issue:', ', '.join(f'{k}={v:.2f}' for v, k in sorted(st, reverse=True)[:8]))
if kre:
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + kre], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(src)))
    hdr = rows[1]
    isrc, isamp, iex = hdr.index('Source'), hdr.index('Warp Stall Sampling (All Samples)'), hdr.index('Instructions Executed')
    data = [(int(r[isamp]), int(r[iex]), r[isrc].strip(), i) for i, r in enumerate(rows[2:]) if len(r) > isamp and r[isamp].isdigit()]
    tot = sum(d[0] for d in data)
    print('total samples', tot)
    for s_, e, t, i in sorted(data, reverse=True)[:n]:
        print(f'{s_:7d} {100*s_/tot:5.1f}% exec={e:9d} #{i:5d} {t[:90]}')
