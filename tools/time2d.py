"""Device-timed 2-D transforms through fft_gpu_plan_2d / fft_gpu_execute (development). usage: time2d.py [rows cols ...]"""
import json, os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import fftb200_loader
F = fftb200_loader.load(); L = F.lib
F.require_gpu()
args = [int(a) for a in sys.argv[1:]] or [1024, 1024, 4096, 4096, 8192, 8192, 16384, 4096, 512, 65536]
for r, c in zip(args[::2], args[1::2]):
    tot = r * c
    mem = L.fft_gpu_alloc(tot)
    L.fftb200_fill_splitmix(L.fftb200_devptr_of(mem), 43, 0, tot)
    plan = L.fft_gpu_plan_2d(r, c, -1)
    for _ in range(3): L.fft_gpu_execute(plan, mem, mem)
    t0 = time.perf_counter()
    reps = 10
    for _ in range(reps): L.fft_gpu_execute(plan, mem, mem)
    ms = (time.perf_counter() - t0) / reps * 1e3
    print(json.dumps({"rows": r, "cols": c, "ms": round(ms, 4), "strict_GBps(32 B/pt)": round(32 * tot / ms * 1e-6), "passes_if_64B/pt": round(64 * tot / ms * 1e-6)}), flush=True)
    L.fft_gpu_destroy_plan(plan); L.fft_gpu_free(mem)
