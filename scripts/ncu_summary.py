"""Text summary of an ncu --set full report (the metrics DESIGN.md / profiles/README.md quote).  usage: ncu_summary.py report.ncu-rep [title]"""
import csv, subprocess, sys
rep = sys.argv[1]; title = sys.argv[2] if len(sys.argv) > 2 else rep
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines())); hdr, units = rows[0], rows[1]
want = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp64.sum", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.sum",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "smsp__thread_inst_executed_pred_on_per_inst_executed.ratio",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
        "l1tex__t_sector_hit_rate.pct", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__inst_executed_op_shared_atom.sum"]
want += [h for h in hdr if "issue_stalled" in h and h.endswith("per_issue_active.ratio")]
print(f"# {title}")
for n, r in enumerate(rows[2:]):
    print(f"## launch {n}: {r[hdr.index('Kernel Name')][:110]}")
    for w in want:
        if w in hdr:
            v = r[hdr.index(w)]
            try:
                v = f"{float(v):.6g}"
            except ValueError:
                pass
            short = w.replace("smsp__average_warps_issue_stalled_", "stall.").replace("_per_issue_active.ratio", "")
            print(f"  {short} = {v} {units[hdr.index(w)]}")
