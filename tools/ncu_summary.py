#!/usr/bin/env python3
"""Summarise an .ncu-rep (raw page + SASS hot spots) - usage: ncu_summary.py report.ncu-rep [top]"""
import csv, subprocess, sys, io
rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 0
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]
want = ['Kernel Name', 'gpu__time_duration.sum', 'launch__registers_per_thread', 'launch__grid_size', 'launch__occupancy_limit',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'smsp__inst_executed.sum', 'sm__inst_executed_pipe_fp64.avg.pct', 'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active', 'sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_lsu.avg.pct',
        'dram__bytes_read.sum', 'dram__bytes_write.sum', 'dram__throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__throughput.avg.pct_of_peak_sustained_elapsed', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum ',
        'smsp__average_warps_issue_stalled', 'sm__cycles_elapsed.max', 'smsp__thread_inst_executed_per_inst_executed.ratio',
        'lts__t_sectors_srcunit_tex_op_red.sum', 'lts__t_sectors_srcunit_tex_op_atom.sum', 'lts__t_sectors_op_red.sum', 'smsp__inst_executed_pipe_fp64', 'local']
for h, u, v in zip(hdr, units, vals):
    if any(h.startswith(w.strip()) for w in want) and not h.endswith(('.max_rate', '.peak_sustained')):
        if 'stalled' in h and not h.endswith('.ratio'):
            continue
        print("%-90s %-12s %s" % (h, u, v))
if top:
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(src)))
    hdr = rows[1]; data = rows[2:]
    iS = hdr.index('Source'); iE = hdr.index('Instructions Executed'); iN = hdr.index('# Samples')
    tot = sum(int(r[iE]) for r in data); totS = sum(int(r[iN]) for r in data)
    print("total warp-instr %d, samples %d, SASS lines %d" % (tot, totS, len(data)))
    if top == 1:
        for k, r in enumerate(data):
            e = int(r[iE]); s = int(r[iN])
            if e > tot * 0.0015 or s > totS * 0.003:
                print("%4d %8.3f %5d  %s" % (k, e / 1e6, s, r[iS].strip()[:100]))
    else:
        for k, r in enumerate(data):
            print("%4d %8.3f %5d  %s" % (k, int(r[iE]) / 1e6, int(r[iN]), r[iS].strip()[:100]))
