"""Condense `ncu --page raw --csv` captures (scripts/ncu_score.sh) into one table, a column per capture:

    python scripts/ncu_summary.py label1=gpurun_out/a.raw.csv label2=gpurun_out/b.raw.csv > profiles/rNN_ncu_score_tc_summary.csv

The LAST column is the one bench.py reads for `roofline.traffic` (dram__bytes_read.sum + dram__bytes_write.sum)."""
import csv, sys
METRICS = [
    "gpu__time_duration.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
    "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
    "l1tex__data_pipe_lsu_wavefronts.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_ld.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_st.sum",
    "l1tex__data_pipe_lsu_wavefronts_mem_lgds.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "l1tex__data_pipe_tc_wavefronts_mem_shared.sum", "l1tex__data_pipe_tc_wavefronts_mem_shared_op_utcmma_matrix_a.sum",
    "l1tex__data_pipe_tc_wavefronts_mem_shared_op_utcmma_matrix_b_scope_1cta.sum",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__sass_inst_executed_op_tmem_ldt.sum", "sm__sass_inst_executed_op_tmem_stt.sum",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum", "lts__t_sectors.sum",
    "lts__t_sector_hit_rate.pct", "smsp__cycles_active.avg", "sm__cycles_elapsed.max", "sm__warps_active.avg.per_cycle_active",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
]
cols = []
for arg in sys.argv[1:]:
    label, path = arg.split("=", 1)
    rows = list(csv.reader(open(path)))
    cols.append((label, {h: (v, u) for h, u, v in zip(rows[0], rows[1], rows[2])}))
w = csv.writer(sys.stdout)
w.writerow(["metric", "unit"] + [l for l, _ in cols])
for m in METRICS:
    unit = next((d[m][1] for _, d in cols if m in d), "")
    w.writerow([m, unit] + [d.get(m, ("", ""))[0] for _, d in cols])
