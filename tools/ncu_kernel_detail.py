"""Prints the issue/stall/pipe picture of one kernel launch of an `ncu --set full` report.
    python tools/ncu_kernel_detail.py report.ncu-rep kernel_substring [occurrence]"""
import csv, io, subprocess, sys
rep, pat = sys.argv[1], sys.argv[2]
occ = int(sys.argv[3]) if len(sys.argv) > 3 else 0
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = rows[0]
sel = [r for r in rows[2:] if pat in r[4]][occ]
want = ["gpu__time_duration.sum", "smsp__inst_executed.sum", "smsp__issue_active.avg.per_cycle_active", "sm__inst_executed_pipe_alu.avg.pct", "sm__inst_executed_pipe_fma.avg.pct",
        "sm__inst_executed_pipe_lsu.avg.pct", "sm__inst_executed_pipe_adu.avg.pct", "sm__inst_executed_pipe_xu.avg.pct", "sm__inst_executed_pipe_cbu.avg.pct", "sm__inst_executed_pipe_uniform.avg.pct",
        "issue_stalled", "sm__warps_active.avg.pct_of_peak", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct", "lts__t_sectors.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "launch__registers_per_thread", "launch__occupancy_limit",
        "smsp__thread_inst_executed_per_inst_executed", "sm__throughput.avg.pct", "l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_st.sum",
        "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum", "lts__t_sectors_op_write.sum", "lts__t_sectors_op_read.sum"]
print(sel[4][:100], sel[8])
for h, u, v in zip(hdr, rows[1], sel):
    if any(k in h for k in want) and v not in ("", "0") and "pcsamp" not in h and ".max" not in h and ".min" not in h:
        print("%-95s %s %s" % (h[-95:], v, u))
