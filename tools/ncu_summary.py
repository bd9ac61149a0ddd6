#!/usr/bin/env python
"""Print the roofline-relevant raw metrics of an ncu report (one line per metric, first kernel)."""
import csv, subprocess, sys, re
rep = sys.argv[1]
rows = list(csv.reader(subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout.splitlines()))
hdr, units = rows[0], rows[1]
keys = ['gpu__time_duration.sum','dram__bytes_read.sum','dram__bytes_write.sum','gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
 'sm__warps_active.avg.pct_of_peak_sustained_active','launch__registers_per_thread','launch__grid_size','launch__block_size','smsp__inst_executed.sum',
 'sm__throughput.avg.pct_of_peak_sustained_elapsed','l1tex__t_sector_hit_rate.pct','lts__t_sector_hit_rate.pct',
 'smsp__issue_active.avg.pct_of_peak_sustained_active','l1tex__m_xbar2l1tex_read_bytes.sum','dram__sectors_read.sum','dram__sectors_write.sum',
 'sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active','sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active',
 'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active',
 'smsp__thread_inst_executed_per_inst_executed.ratio','sm__cycles_elapsed.avg','lts__t_sectors_srcunit_tex_op_write.sum','lts__t_sectors_srcunit_tex_op_read.sum',
 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum','smsp__warps_eligible.avg.per_cycle_active','smsp__warps_active.avg.per_cycle_active']
for r in rows[2:3]:
    print('kernel', r[hdr.index('Kernel Name')])
    for k in keys:
        if k in hdr: print(f"{k:75s} {r[hdr.index(k)]:>16s} {units[hdr.index(k)]}")
    print('--- stall reasons (warps per issue-active cycle)')
    for i,h in enumerate(hdr):
        m = re.match(r'smsp__average_warps_issue_stalled_(.*)_per_issue_active.ratio', h)
        if m:
            try: v=float(r[i])
            except: continue
            if v > 0.05: print(f"   {m.group(1):30s} {v:.3f}")
