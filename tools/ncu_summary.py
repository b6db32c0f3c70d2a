"""Summarises one `ncu --page raw --csv` dump: headline metrics, warp stall reasons, pipe utilisation."""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr, units = rows[0], rows[1]
for vals in rows[2:]:
    d = {h: (u, v) for h, u, v in zip(hdr, units, vals)}
    print("== kernel:", d.get("Kernel Name", ("", "?"))[1], "grid", d.get("Grid Size", ("", "?"))[1], "block", d.get("Block Size", ("", "?"))[1])
    want = ['gpu__time_duration.sum', 'launch__registers_per_thread', 'launch__occupancy_limit_registers',
            'launch__occupancy_limit_shared_mem', 'launch__occupancy_limit_warps', 'launch__waves_per_multiprocessor',
            'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
            'smsp__inst_executed.sum', 'sm__inst_executed.avg.per_cycle_elapsed', 'smsp__inst_executed.avg.per_cycle_active',
            'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'smsp__thread_inst_executed_per_inst_executed.ratio',
            'smsp__sass_thread_inst_executed_op_ffma_pred_on.sum', 'smsp__sass_thread_inst_executed_op_fadd_pred_on.sum',
            'smsp__sass_thread_inst_executed_op_fmul_pred_on.sum', 'smsp__sass_thread_inst_executed_op_dfma_pred_on.sum',
            'smsp__sass_thread_inst_executed_op_dadd_pred_on.sum', 'smsp__sass_thread_inst_executed_op_dmul_pred_on.sum',
            'lts__t_bytes.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'smsp__sass_inst_executed_op_local_ld.sum',
            'smsp__sass_inst_executed_op_local_st.sum', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
            'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'smsp__sass_inst_executed_op_shared_ld.sum']
    for w in want:
        if w in d:
            print(f"  {w:70s} {d[w][1]:>18s} {d[w][0]}")
    print("  -- warp stall reasons (warps per issue-active cycle, > 0.05)")
    st = []
    for k in d:
        if 'smsp__average_warp' in k and 'issue_stalled' in k and k.endswith('.ratio') and 'not_issued' not in k:
            try:
                st.append((float(d[k][1]), k.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', '')))
            except ValueError:
                pass
    for v, k in sorted(st, reverse=True):
        if v > 0.05:
            print(f"     {k:40s} {v:.3f}")
    print("  -- pipe utilisation (% of peak sustained active, > 1)")
    pp = []
    for k in d:
        if k.startswith('sm__inst_executed_pipe_') and k.endswith('.avg.pct_of_peak_sustained_active'):
            try:
                pp.append((float(d[k][1]), k.replace('sm__inst_executed_pipe_', '').replace('.avg.pct_of_peak_sustained_active', '')))
            except ValueError:
                pass
    for v, k in sorted(pp, reverse=True):
        if v > 1:
            print(f"     {k:40s} {v:.2f}")
