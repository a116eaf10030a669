#!/usr/bin/env python
"""Summarise one ncu --set full report: python tools_ncu_summary.py gpurun_out/x.ncu-rep [out.txt]"""
import csv, subprocess, sys
rep = sys.argv[1]
out = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]
want = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'dram__throughput.avg.pct_of_peak_sustained_elapsed', 'lts__t_sector_hit_rate.pct', 'l1tex__t_sector_hit_rate.pct',
        'l1tex__t_sectors_lookup_hit.sum', 'l1tex__t_sectors_lookup_miss.sum', 'l1tex__m_xbar2l1tex_read_bytes.sum',
        'lts__t_sectors_srcunit_tex_op_read.sum', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__data_pipe_lsu_wavefronts.sum',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_ld.sum', 'smsp__inst_executed_op_shared_ld.sum', 'sm__inst_executed_pipe_lsu.sum',
        'smsp__warp_issue_stalled_membar_per_warp_active.pct', 'smsp__warp_issue_stalled_sleeping_per_warp_active.pct', 'l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum',
        'l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum', 'l1tex__lsu_writeback_active.avg.pct_of_peak_sustained_elapsed',
        'l1tex__lsuin_requests.avg.pct_of_peak_sustained_elapsed', 'l1tex__data_bank_reads.avg.pct_of_peak_sustained_elapsed',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum', 'sm__inst_executed_pipe_lsu.sum',
        'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_fp64.sum',
        'sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active', 'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active',
        'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size', 'launch__occupancy_limit_registers',
        'launch__occupancy_limit_shared_mem', 'sm__maximum_warps_per_active_cycle_pct', 'smsp__cycles_active.avg',
        'smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct', 'smsp__warp_issue_stalled_lg_throttle_per_warp_active.pct',
        'smsp__warp_issue_stalled_short_scoreboard_per_warp_active.pct', 'smsp__warp_issue_stalled_math_pipe_throttle_per_warp_active.pct',
        'smsp__warp_issue_stalled_wait_per_warp_active.pct', 'smsp__warp_issue_stalled_no_instruction_per_warp_active.pct',
        'smsp__warp_issue_stalled_mio_throttle_per_warp_active.pct', 'smsp__warp_issue_stalled_barrier_per_warp_active.pct',
        'smsp__warp_issue_stalled_not_selected_per_warp_active.pct', 'smsp__warp_issue_stalled_dispatch_stall_per_warp_active.pct',
        'smsp__warp_issue_stalled_tex_throttle_per_warp_active.pct', 'smsp__warp_issue_stalled_branch_resolving_per_warp_active.pct',
        'smsp__thread_inst_executed_per_inst_executed.ratio'] + [
        f'smsp__average_warps_issue_stalled_{r}_per_issue_active.ratio' for r in (
            'long_scoreboard', 'short_scoreboard', 'mio_throttle', 'lg_throttle', 'math_pipe_throttle', 'wait', 'not_selected',
            'barrier', 'branch_resolving', 'dispatch_stall', 'no_instruction', 'tex_throttle', 'sleeping', 'selected')] + [
        'lts__t_sectors_srcunit_tex_lookup_hit.sum', 'lts__t_sectors_srcunit_tex_lookup_miss.sum',
        'lts__average_t_sector_hit_rate_srcunit_tex_realtime.pct', 'l1tex__average_t_sector_hit_rate_realtime.pct']
lines = []
for r in rows[2:]:
    name = r[hdr.index('Kernel Name')] if 'Kernel Name' in hdr else '?'
    lines.append(f'== {name[:100]}')
    for w in want:
        if w in hdr:
            i = hdr.index(w)
            lines.append(f'{w:75s} {r[i]:>20s} {units[i]}')
txt = '\n'.join(lines)
print(txt)
if len(sys.argv) > 2:
    open(sys.argv[2], 'w').write(f'# ncu --set full summary of {rep}\n' + txt + '\n')
