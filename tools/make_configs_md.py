"""gpurun_out/configs_bench.log (tools/configs_bench.py) -> profiles/<tag>_configs.md. usage: make_configs_md.py <tag> [peak GB/s]"""
import json, os, sys
tag = sys.argv[1]; peak = float(sys.argv[2]) if len(sys.argv) > 2 else 6650.0
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
rows = []
for l in open(os.path.join(ROOT, "gpurun_out", "configs_bench.log")):
    try: rows.append(json.loads(l))
    except ValueError: pass
with open(os.path.join(ROOT, "profiles", tag + "_configs.md"), "w") as f:
    f.write("# %s - every BASELINE configuration that fits one GPU, device-timed through the engine C-ABI (tools/configs_bench.py)\n\n" % tag)
    f.write("CUDA events on the plan's stream, best of 10 after 3 warm-ups, one B200 (gpurun). `strict` = algorithmic bytes (32 B per point: 16 read + 16 written) / time;\n"
            "roofline = strict / %.0f GB/s (B200_PROFILING.md fallback peak; MEASURED_PEAKS.json was absent). SM clocks differ by up to 7 %% between boxes of the pool (power capping):\n"
            "the compute-bound kernels (N = 1024, 2048, 8192) move with them, the HBM-bound ones do not.\n\n" % peak)
    f.write("| config | n | batch | ms best | ms median | TFLOP/s (5 N log2 N) | strict GB/s | frac of HBM roofline | plan |\n|---|---|---|---|---|---|---|---|---|\n")
    for r in rows:
        if "n" in r and "strict_GBps" in r:
            f.write("| %s | %d | %d | %.4f | %.4f | %.2f | %d | %.2f | %s |\n" % (r["cfg"], r["n"], r["batch"], r["ms_best"], r["ms_med"], r["gflops"] / 1e3, r["strict_GBps"],
                                                                           r["strict_GBps"] / peak, r["plan"].split(": ", 1)[-1]))
        else:
            rest = {k: v for k, v in r.items() if k not in ("cfg", "ms_best", "plan")}
            f.write("| %s | | | %s | | | | | %s %s |\n" % (r["cfg"], r.get("ms_best", r.get("us_per_call", "")) if "ms_best" in r else "%.1f us" % r["us_per_call"], json.dumps(rest), r.get("plan", "").split(": ", 1)[-1]))
print(open(os.path.join(ROOT, "profiles", tag + "_configs.md")).read())
