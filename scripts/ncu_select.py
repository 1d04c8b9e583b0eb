"""Selected metrics of one `ncu --set full` capture (the `--page raw --csv` export: header row, unit row, one row per launch) in
the three-column form kept under profiles/ (metric,value,unit).
    python scripts/ncu_select.py gpurun_out/r02_rollout_tc2_full_raw.csv > profiles/r02_rollout_tc2_ncu_full_selected.csv"""
import csv
import sys

KEYS = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__cluster_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_sector_hit_rate.pct",
        "l1tex__m_xbar2l1tex_read_bytes.sum", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__cycles_active.avg", "sm__cycles_elapsed.max", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_st.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_ld.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smsp__inst_executed.sum", "sm__warps_active.avg.pct_of_peak_sustained_active"]

rows = list(csv.reader(open(sys.argv[1])))
head, units, vals = rows[0], rows[1], rows[-1]
col = {name: i for i, name in enumerate(head)}
out = csv.writer(sys.stdout, lineterminator="\n")
out.writerow(["metric", "value", "unit"])
out.writerow(["kernel", vals[col["Kernel Name"]], ""])
for k in KEYS:
    if k in col:
        out.writerow([k, vals[col[k]], units[col[k]]])
