#!/usr/bin/env python
"""Turn an .ncu-rep (read on the CPU box with `ncu -i`) into the small text summary that is committed under profiles/.
Usage: python profiles/ncu_summary.py gpurun_out/prof.ncu-rep > profiles/r01_k_match_iquv.txt"""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]
want = ['Kernel Name', 'gpu__time_duration.sum', 'launch__grid_size', 'launch__block_size', 'launch__registers_per_thread',
        'launch__shared_mem_per_block_dynamic', 'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active', 'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_tma.avg.pct_of_peak_sustained_active', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'dram__bytes_read.sum', 'dram__bytes_write.sum', 'dram__throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__t_bytes.sum', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'sm__cycles_elapsed.avg.per_second',
        'smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio',
        'smsp__thread_inst_executed_per_inst_executed.ratio']
print('# ncu --set full --clock-control none, report', rep.split('/')[-1])
for w in want:
    for h, u, v in zip(hdr, units, vals):
        if h == w:
            print('%-90s %-12s %s' % (h, u, v))
src = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
h2 = rows[1]
ix = {h: i for i, h in enumerate(h2)}


def num(x):
    try:
        return float(x)
    except ValueError:
        return 0.0


data = [r for r in rows[2:] if len(r) == len(h2)]
tot = sum(num(r[ix['Instructions Executed']]) for r in data)
ts = sum(num(r[ix['# Samples']]) for r in data)
ops = {}
for r in data:
    t = r[ix['Source']].split()
    op = t[1] if t and t[0].startswith('@') else (t[0] if t else '')
    o = ops.setdefault(op, [0, 0])
    o[0] += num(r[ix['Instructions Executed']])
    o[1] += num(r[ix['# Samples']])
print('\n# executed warp instructions by opcode (share of %.3e), warp-stall samples (share of %d)' % (tot, ts))
for op, (i, s) in sorted(ops.items(), key=lambda kv: -kv[1][0])[:18]:
    print('%-24s inst %6.2f%%   samples %6.2f%%' % (op, i / tot * 100, s / ts * 100))
mx = max(num(r[ix['Instructions Executed']]) for r in data if 'FFMA' in r[ix['Source']])
hot = [r for r in data if num(r[ix['Instructions Executed']]) >= 0.99 * mx]
print('\n# hot loop: %d SASS instructions, each executed %.0f times: %.1f%% of executed instructions, %.1f%% of samples; FFMA share inside: %.1f%%'
      % (len(hot), mx, sum(num(r[ix['Instructions Executed']]) for r in hot) / tot * 100, sum(num(r[ix['# Samples']]) for r in hot) / ts * 100,
         sum(1 for r in hot if 'FFMA' in r[ix['Source']]) / len(hot) * 100))
