#!/usr/bin/env python
"""Print the metrics we track from an .ncu-rep (raw page) -- used to write profiles/*.md."""
import csv, subprocess, sys
WANT = ['gpu__time_duration.sum','dram__bytes_read.sum','dram__bytes_write.sum','dram__throughput.avg.pct_of_peak_sustained_elapsed',
 'lts__t_sector_hit_rate.pct','lts__t_sectors_srcunit_tex_op_read.sum','lts__t_sectors_srcunit_tex_lookup_hit.sum','lts__t_sectors_srcunit_tex_lookup_miss.sum',
 'lts__t_sectors_srcunit_tex_evict_first_lookup_miss.sum','lts__t_sectors_srcunit_tex_evict_normal_lookup_hit.sum','lts__t_sectors_srcunit_tex_evict_normal_lookup_miss.sum',
 'lts__throughput.avg.pct_of_peak_sustained_elapsed','l1tex__t_sector_hit_rate.pct','l1tex__throughput.avg.pct_of_peak_sustained_elapsed',
 'l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum','l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum','l1tex__data_pipe_lsu_wavefronts.sum','l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
 'sm__warps_active.avg.pct_of_peak_sustained_active','launch__registers_per_thread','launch__occupancy_limit_registers','launch__grid_size','smsp__inst_executed.sum',
 'sm__cycles_elapsed.max','smsp__issue_active.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
 'smsp__pcsamp_warps_issue_stalled_long_scoreboard','smsp__pcsamp_warps_issue_stalled_lg_throttle','smsp__pcsamp_warps_issue_stalled_short_scoreboard','smsp__pcsamp_warps_issue_stalled_math_pipe_throttle',
 'smsp__pcsamp_warps_issue_stalled_wait','smsp__pcsamp_warps_issue_stalled_not_selected','smsp__pcsamp_warps_issue_stalled_selected','smsp__pcsamp_warps_issue_stalled_barrier',
 'smsp__pcsamp_warps_issue_stalled_mio_throttle','smsp__pcsamp_warps_issue_stalled_branch_resolving','smsp__pcsamp_warps_issue_stalled_dispatch_stall','smsp__pcsamp_warps_issue_stalled_no_instructions',
 'l1tex__m_xbar2l1tex_read_sectors.sum','l1tex__m_l1tex2xbar_req_cycles_active.avg.pct_of_peak_sustained_elapsed','lts__t_sectors_srcunit_ltcfabric.sum']
for path in sys.argv[1:]:
    out = subprocess.run(['ncu','-i',path,'--page','raw','--csv'],capture_output=True,text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    for vals in rows[2:]:
        print(f"=== {path}: {vals[hdr.index('Kernel Name')][:60]}")
        for w in WANT:
            if w in hdr:
                i = hdr.index(w); print(f"{w:78s} {vals[i]:>22s} {units[i]}")
