#!/usr/bin/env python
"""The handful of raw metrics that decide what bounds k_raster (issue slots, the LSU data pipe, barriers), for every launch
of an ncu report.  usage: tools/ncu_key.py report.ncu-rep"""
import csv, io, subprocess, sys
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out))); hdr, units = rows[0], rows[1]
want = ['gpu__time_duration.sum', 'Grid Size', 'launch__registers_per_thread', 'smsp__inst_executed.sum', 'smsp__thread_inst_executed_per_inst_executed.ratio',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__warps_eligible.avg.per_cycle_active',
        'l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed', 'l1tex__data_pipe_lsu_wavefronts_mem_shared_op_ld.sum',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared_op_st.sum', 'l1tex__data_pipe_lsu_wavefronts_mem_shared_op_atom.sum',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'l1tex__t_output_wavefronts_pipe_lsu_mem_global_op_ld.sum',
        'l1tex__lsu_writeback_active_mem_lgds.sum.pct_of_peak_sustained_elapsed', 'sm__inst_executed_pipe_lsu.sum', 'sm__inst_executed_pipe_alu.sum',
        'sm__inst_executed_pipe_fma.sum', 'sm__inst_executed_pipe_xu.sum', 'sm__inst_executed_pipe_uniform.sum', 'sm__inst_executed_pipe_cbu.sum', 'sm__inst_executed_pipe_adu.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'sm__cycles_active.avg',
        'smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_membar_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_selected_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_sleeping_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_imc_miss_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_tex_throttle_per_issue_active.ratio']
for r in rows[2:]:
    print("---- " + r[hdr.index('Kernel Name')][:60])
    for i, h in enumerate(hdr):
        if h in want:
            print(f"{h:92s} {r[i]:>16s} {units[i]}")
