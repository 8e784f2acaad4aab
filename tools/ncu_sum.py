"""One-screen summary of an ncu report: usage: python tools/ncu_sum.py X.ncu-rep"""
import csv, io, re, subprocess, sys
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
want = [r"^Kernel Name$", r"gpu__time_duration.sum", r"dram__bytes_read.sum$", r"dram__bytes_write.sum$", r"dram__throughput.avg.pct_of_peak_sustained_elapsed",
        r"sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", r"l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
        r"l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct", r"l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum$", r"l1tex__data_pipe_lsu_wavefronts_mem_shared.sum$",
        r"sm__warps_active.avg.pct_of_peak", r"launch__registers_per_thread", r"smsp__average_warps_issue_stalled_.*_per_issue_active.ratio",
        r"sm__cycles_elapsed.avg.per_second", r"smsp__inst_executed.sum$", r"lts__t_sector_hit_rate.pct", r"sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active"]
for k in range(2, len(rows)):
    print("----")
    for h, u, v in zip(rows[0], rows[1], rows[k]):
        if any(re.search(w, h) for w in want):
            if "issue_stalled" in h:
                try:
                    if float(v) < 0.3: continue
                except ValueError: pass
                h = h.replace("smsp__average_warps_issue_stalled_", "stall ").replace("_per_issue_active.ratio", "")
            print("%-75s %-8s %s" % (h, u, v))
