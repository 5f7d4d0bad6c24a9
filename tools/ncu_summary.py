"""Key metrics of every kernel of one ncu report: python tools/ncu_summary.py report.ncu-rep"""
import csv
import subprocess
import sys

raw = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
col = {h: i for i, h in enumerate(hdr)}
want = ["Kernel Name", "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed.avg.per_cycle_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "smsp__warps_active.avg.per_cycle_active", "smsp__warps_eligible.avg.per_cycle_active",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "sass__inst_executed_local_loads", "sass__inst_executed_local_stores"]
for vals in rows[2:]:
    if len(vals) < len(hdr):
        continue
    for w in want:
        if w in col:
            print(f"{w:75s} {vals[col[w]]:>16s} {units[col[w]]}")
    st = []
    for h, i in col.items():
        if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("_per_issue_active.ratio"):
            try:
                st.append((float(vals[i].replace(",", "")), h[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")]))
            except ValueError:
                pass
    print("stall cycles per issued instruction:", ", ".join(f"{n} {v:.2f}" for v, n in sorted(st, reverse=True)[:8]))
