"""One line per kernel of an ncu report with the counters that matter here:  python tools/ncu_kernels.py REPORT.ncu-rep"""
import csv, io, subprocess, sys
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
h = rows[0]
cols = [("ms", "gpu__time_duration.sum"), ("winst_G", "smsp__inst_executed.sum"), ("lanes", "smsp__thread_inst_executed_per_inst_executed.ratio"),
        ("issue%", "sm__issue_active.avg.pct_of_peak_sustained_elapsed"), ("lsu_wf%", "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed"),
        ("warps%", "sm__warps_active.avg.pct_of_peak_sustained_active"), ("L1hit", "l1tex__t_sector_hit_rate.pct"), ("L2hit", "lts__t_sector_hit_rate.pct"),
        ("dramR_GB", "dram__bytes_read.sum"), ("dramW_GB", "dram__bytes_write.sum"),
        ("long_sb", "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio"),
        ("short_sb", "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio"),
        ("wait", "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio"),
        ("math", "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio"),
        ("branch", "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio"),
        ("regs", "launch__registers_per_thread"), ("grid", "launch__grid_size")]
for r in rows[2:]:
    name = r[h.index("Kernel Name")].split("(")[0].replace("void ", "")
    vals = []
    for k, c in cols:
        if c in h:
            try:
                x = float(r[h.index(c)].replace(",", ""))
                if k == "winst_G": x /= 1e9
                vals.append(f"{k}={x:.2f}" if x < 1000 else f"{k}={x:.0f}")
            except ValueError:
                pass
    print(name, " ".join(vals))
