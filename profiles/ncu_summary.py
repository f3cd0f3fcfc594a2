#!/usr/bin/env python
"""Condense an .ncu-rep (ncu --set full) into the handful of counters this project tracks.
usage: python profiles/ncu_summary.py gpurun_out/prof.ncu-rep [out.json]"""
import csv, io, json, subprocess, sys
KEEP = ['Kernel Name', 'Grid Size', 'Block Size', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'launch__registers_per_thread', 'launch__shared_mem_per_block_dynamic', 'launch__occupancy_limit_registers',
        'launch__occupancy_limit_shared_mem', 'launch__occupancy_limit_warps', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'smsp__inst_executed.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
        'lts__t_sectors_srcunit_tex_op_red.sum', 'lts__t_sectors_srcunit_tex_op_read.sum',
        'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__cycles_elapsed.max', 'smsp__inst_executed_op_local_ld.sum', 'smsp__inst_executed_op_local_st.sum',
        'smsp__inst_executed_op_shared_ld.sum', 'smsp__inst_executed_op_shared_st.sum', 'smsp__inst_executed_op_global_ld.sum',
        'smsp__inst_executed_op_global_red.sum', 'l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum',
        'l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum', 'smsp__thread_inst_executed_per_inst_executed.ratio']
STALL = 'smsp__average_warps_issue_stalled_'
def main():
    rep = sys.argv[1]
    out = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], stdout=subprocess.PIPE, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    res = []
    for vals in rows[2:]:
        d = {}
        for h, u, v in zip(hdr, units, vals):
            if h in KEEP:
                d[h] = f"{v} {u}".strip()
            elif h.startswith(STALL) and h.endswith('_per_issue_active.ratio'):
                d.setdefault('stalls_per_issue', {})[h[len(STALL):-len('_per_issue_active.ratio')]] = round(float(v), 3)
        res.append(d)
    txt = json.dumps(res, indent=1)
    if len(sys.argv) > 2:
        open(sys.argv[2], 'w').write(txt)
    print(txt)
main()
