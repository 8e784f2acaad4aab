"""gpurun_out/ (tools/refresh_profiles_r02.sh) -> the tracked round-2 summaries under profiles/. usage: python tools/make_r02_profiles.py"""
import csv, io, json, os, re, shutil, subprocess, sys
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")

# 1. bench lines, cuFFT table
for f in ("r02_bench_line.json", "r02_bench_reference_line.json"):
    line = open(os.path.join(G, f)).read().strip().splitlines()[-1]
    json.dump(json.loads(line), open(os.path.join(P, f), "w"), indent=1)
shutil.copy(os.path.join(G, "r02_cufft.md"), os.path.join(P, "r02_cufft.md"))
d = json.load(open(os.path.join(P, "r02_bench_line.json")))
ref = json.load(open(os.path.join(P, "r02_bench_reference_line.json")))

# 2. configs table from the secondary block
peak = d["roofline"]["peak"]
with open(os.path.join(P, "r02_configs.md"), "w") as f:
    f.write("# r02 - every BASELINE configuration that fits one GPU, from the `secondary` block of the bench line itself (profiles/r02_bench_line.json)\n\n")
    f.write("Same run as the headline (python bench.py on one B200, gpurun): CUDA events on the plan's stream, 3 warm-ups + 12 executions timed one by one (`ms` = median, `best` = minimum); "
            "`frac` = strict GB/s of the median / %.1f (MEASURED_PEAKS.json hbm_gbs - a copy with a fixed share per SM; copies whose tiles are taken on demand reach 6.9-7.1 TB/s on this chip, "
            "profiles/r02_microbench.md section 5, which is why the HBM-bound sizes read above 1);\n" % peak)
    c = d["clocks"]
    f.write("`rel L2` = sampled transform(s) against the CPU oracle. SM clock during the run: %.1f MHz of %.1f (reasons: %s) - the compute-bound sizes (2^13 and up) move with it.\n\n"
            % (c["sm_mhz"], c["sm_max_mhz"], ", ".join(c["reasons"]) or "none"))
    f.write("Headline (cfg2, N=4096 x 65536): %.4f ms/step, %.0f GFLOP/s, roofline frac %.3f; e2e %.1f GFLOP/s (%.1f ms/step, 4 GiB each way); reference CPU arm %.1f GFLOP/s on %d cores "
            "-> e2e ratio %.2fx, device ratio %.0fx.\n\n" % (d["ms_per_step"], d["value"], d["roofline"]["frac"], d["e2e"]["value"], d["e2e"]["ms_per_step"], ref["value"],
                                                          ref["cpu_baseline"]["cores"], d["e2e"]["value"] / ref["value"], d["value"] / ref["value"]))
    f.write("(N = 4096 appears twice: the headline times its executions back to back, the band row one by one with an event pair and a synchronise around each - on the "
            "power-capped boxes of this pool the one-by-one median of this size, the one with the most arithmetic per byte among the single-visit kernels, sits 10-15 %% above its own best.)\n\n")
    f.write("| config | n | batch | ms (median) | best | TFLOP/s | strict GB/s | frac of HBM roofline | launches | rel L2 vs oracle | plan |\n|---|---|---|---|---|---|---|---|---|---|---|\n")
    for s in d["secondary"]:
        f.write("| %s | %d | %d | %.4f | %.4f | %.2f | %.0f | %.3f | %d | %.1e | %s |\n" % (s["config"], s["n"], s["batch"], s["ms"], s["ms_best"], s["gflops"] / 1e3, s["strict_GBps"], s["frac"],
                                                                                             s["launches"], s["rel_l2_vs_oracle"], s["plan"].split(": ", 1)[-1]))
    o = d["cpu_baseline"].get("others", [])
    if o:
        f.write("\nCPU baselines of the same run (`cpu_baseline.others`): " + "; ".join("%s N=%d: %.2f GFLOP/s (%d core%s)" % (x["name"].split(" (")[0], x["n"], x["value"], x["cores"], "s" if x["cores"] > 1 else "") for x in o) + "\n")

# 3. launch list + full capture of the headline kernel (tools/make_profiles.py), one-screen summaries of the others (tools/ncu_sum.py)
subprocess.check_call([sys.executable, os.path.join(ROOT, "tools", "make_profiles.py"), "r02", os.path.join(G, "r02_launches.csv"), os.path.join(G, "r02_pipe.ncu-rep"), "n4096_b65536"])
traffic = json.load(open(os.path.join(P, "traffic.json")))
for rep, key in (("r02_pipe13", "n8192_b32768_pipe13"), ("r02_fused_16", "n65536_b4096_fused"), ("r02_fused_20", "n1048576_b256_fused")):
    path = os.path.join(G, rep + ".ncu-rep")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_sum.py"), path], capture_output=True, text=True).stdout
    out = "\n".join(l for l in out.splitlines() if l.strip() != "----")
    with open(os.path.join(P, "r02_%s_ncu.md" % key), "w") as f:
        f.write("# r02 - ncu --set full, one launch, --clock-control none: %s (tools/ncu_sum.py gpurun_out/%s.ncu-rep; report is scratch, not tracked)\n\n```\n%s\n```\n" % (key, rep, out))
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    r = list(csv.reader(io.StringIO(raw)))
    m = {h: (u, v) for h, u, v in zip(r[0], r[1], r[2])}
    sc = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}
    traffic[key] = sum(float(m[k][1].replace(",", "")) * sc[m[k][0]] for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"))
json.dump(traffic, open(os.path.join(P, "traffic.json"), "w"), indent=1, sort_keys=True)
print(open(os.path.join(P, "r02_configs.md")).read()[:1500])
