"""Turns the raw ncu outputs in gpurun_out/ into the tracked summaries under profiles/.
usage: python tools/make_profiles.py <round tag, e.g. r01> <launch csv> <ncu-rep> <key for traffic.json>"""
import collections, csv, io, json, os, re, subprocess, sys
tag, launches, rep, key = sys.argv[1:5]
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
P = os.path.join(ROOT, "profiles")

rows = [r for r in csv.reader(open(launches)) if len(r) > 5]
ix = {h: i for i, h in enumerate(rows[0])}
groups = collections.OrderedDict()
for r in rows[1:]:
    try:
        ns = float(r[ix["Metric Value"]].replace(",", ""))
    except ValueError:
        continue
    name = re.sub(r"\(.*", "", r[ix["Kernel Name"]]).strip()
    if "fft_" in name and "fill" not in name:
        name += " [whole batch]" if ns > 3e5 else " [e2e chunk]"
    k = (name, r[ix["Grid Size"]], r[ix["Block Size"]])
    g = groups.setdefault(k, [])
    g.append(ns)
total = sum(sum(v) for v in groups.values())
with open(os.path.join(P, tag + "_launches.md"), "w") as f:
    f.write("# %s - ncu launch list (gpu__time_duration.sum, --clock-control none)\n\n" % tag)
    f.write("Command: see the header of gpurun_out/%s (bench.py --steps 2 --warmup 3 --no-cpu --e2e-steps 1).\n" % os.path.basename(launches))
    f.write("Per-launch times under ncu are cold-cache and serialised: compare shares, not absolutes.\n\n")
    f.write("| kernel | grid | block | launches | total us | share | avg us | min us |\n|---|---|---|---|---|---|---|---|\n")
    for (name, grid, block), v in groups.items():
        f.write("| %s | %s | %s | %d | %.1f | %.1f%% | %.1f | %.1f |\n" % (name, grid, block, len(v), sum(v) / 1e3, 100 * sum(v) / total, sum(v) / len(v) / 1e3, min(v) / 1e3))
    f.write("\n[whole batch] launches are the timed step of bench.py (one launch per step: 3 warm-up + 1 + 2 timed + clock-sampling tail).\n"
            "[e2e chunk] launches belong to the e2e leg (fft_gpu_dft_1d_batch: 128 chunks of 32 MiB per call, overlapped with PCIe copies).\n")

out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
r = list(csv.reader(io.StringIO(out)))
m = {h: (u, v) for h, u, v in zip(r[0], r[1], r[2])}
def val(k):
    return float(m[k][1].replace(",", ""))
def scale(k):  # to bytes
    return {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}[m[k][0]]
rd, wr = val("dram__bytes_read.sum") * scale("dram__bytes_read.sum"), val("dram__bytes_write.sum") * scale("dram__bytes_write.sum")
keys = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "launch__grid_size", "launch__block_size",
        "sm__cycles_elapsed.avg.per_second", "smsp__inst_executed.sum", "lts__t_sector_hit_rate.pct",
        "smsp__sass_thread_inst_executed_op_dfma_pred_on.sum", "smsp__sass_thread_inst_executed_op_dadd_pred_on.sum",
        "smsp__sass_thread_inst_executed_op_dmul_pred_on.sum"]
with open(os.path.join(P, tag + "_" + key + "_ncu.md"), "w") as f:
    f.write("# %s - ncu --set full of the dominant kernel (%s)\n\n" % (tag, r[2][r[0].index("Kernel Name")]))
    f.write("Source report: gpurun_out/%s (scratch, not tracked). One launch, --clock-control none.\n\n| metric | unit | value |\n|---|---|---|\n" % os.path.basename(rep))
    for k in keys:
        if k in m:
            f.write("| %s | %s | %s |\n" % (k, m[k][0], m[k][1]))
    for k in sorted(m):
        if "issue_stalled" in k and k.endswith("_per_issue_active.ratio"):
            try:
                if float(m[k][1]) >= 0.3:
                    f.write("| %s | %s | %s |\n" % (k, m[k][0], m[k][1]))
            except ValueError:
                pass
    f.write("\nDRAM traffic per launch: %.3f GB read + %.3f GB written = %.3f GB.\n" % (rd / 1e9, wr / 1e9, (rd + wr) / 1e9))
tj = os.path.join(P, "traffic.json")
d = json.load(open(tj)) if os.path.exists(tj) else {}
d[key] = rd + wr
json.dump(d, open(tj, "w"), indent=1, sort_keys=True)
print(open(os.path.join(P, tag + "_launches.md")).read())
print(open(os.path.join(P, tag + "_" + key + "_ncu.md")).read())
