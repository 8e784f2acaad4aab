"""Development check on a B200 (run under gpurun): parity sweep vs the oracle + quick timings."""
import ctypes as C, json, math, os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import numpy as np
import fftb200_loader
from oracle import oracle as O

F = fftb200_loader.load(); L = F.lib
F.require_gpu()
p = O.port()
print("device:", L.fft_gpu_get_device_name().decode(), "SMs", L.fftb200_sm_count(), flush=True)

def check(n, batch, seed=43):
    x = p.fill(seed, 0, n * batch).reshape(batch, n)
    res = {}
    for d in (-1, 1):
        try:
            y = F.gpu_fft_batch(x, d)
        except Exception as e:
            return {"n": n, "batch": batch, "error": str(e)}
        if n & (n - 1) == 0:
            ref = p.fft_batch(x, d)
        else:
            ref = np.stack([p.fft(r, d) for r in x])
        res["fwd" if d < 0 else "inv"] = O.rel_l2(y, ref)
        if d < 0:
            res["vs_numpy"] = O.rel_l2(y, np.fft.fft(x, axis=1))
            yi = F.gpu_fft_batch(y, 1, inplace=True)
            res["roundtrip"] = O.rel_l2(yi, x)
    return {"n": n, "batch": batch, **res}

sizes = [(1 << l, max(1, min(64, (1 << 16) >> l))) for l in range(1, 21)]
if len(sys.argv) > 1 and sys.argv[1] == "big":
    sizes = [(1 << l, 1) for l in range(21, 25)]
for n, b in sizes:
    print(json.dumps(check(n, b + (3 if n <= 4096 else 0))), flush=True)
for n, b in [(97, 3), (1009, 2), (100003, 1), (1000003, 1), (6, 5), (12, 2)]:
    print(json.dumps(check(n, b, 46)), flush=True)

# r2c
for n in (1024, 1 << 14, 1 << 20):
    xr = p.fill(47, 0, n).real.copy()
    print(json.dumps({"r2c": n, "err": O.rel_l2(F.r2c(xr), p.r2c(xr))}), flush=True)
# fft_auto host path
x = p.fill(42, 0, 1024)
print(json.dumps({"fft_auto_1024": O.rel_l2(F.fft_auto(x), p.fft(x))}), flush=True)

def timeit(n, batch, reps=10, inplace=False):
    tot = n * batch
    m_in = L.fft_gpu_alloc(tot); m_out = m_in if inplace else L.fft_gpu_alloc(tot)
    L.fftb200_fill_splitmix(L.fftb200_devptr_of(m_in), 43, 0, tot)
    plan = L.fft_gpu_plan_1d(n, batch, -1)
    eng = L.fftb200_engine_of(plan)
    din, dout = L.fftb200_devptr_of(m_in), L.fftb200_devptr_of(m_out)
    for _ in range(3): L.fftb200_plan_exec(eng, din, dout)
    ts = []
    ms = C.c_float()
    for _ in range(reps):
        L.fftb200_timer_start(eng); L.fftb200_plan_exec_async(eng, din, dout); L.fftb200_timer_stop(eng, C.byref(ms)); ts.append(ms.value)
    desc = L.fftb200_plan_describe(eng).decode()
    L.fft_gpu_destroy_plan(plan); L.fft_gpu_free(m_in)
    if not inplace: L.fft_gpu_free(m_out)
    t = min(ts)
    return {"n": n, "batch": batch, "ms_best": t, "ms_med": sorted(ts)[len(ts)//2], "gflops": 5*n*math.log2(n)*batch/t*1e-6,
            "strict_GBps": 32*n*batch/t*1e-6, "plan": desc}

for lg in (12, 10, 8, 6, 11, 13, 14, 16, 17, 18, 20, 22, 24):
    print(json.dumps(timeit(1 << lg, (1 << 28) >> lg)), flush=True)

try:
    import torch
    def bench(n, batch, reps=10):
        x = torch.randn(batch, n, dtype=torch.complex128, device="cuda"); y = torch.fft.fft(x); torch.cuda.synchronize(); ts = []
        for _ in range(reps):
            a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
            a.record(); y = torch.fft.fft(x); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
        t = min(ts); print(json.dumps({"cufft": 1, "n": n, "batch": batch, "ms": t, "gflops": 5*n*math.log2(n)*batch/t*1e-6, "strict_GBps": 32*n*batch/t*1e-6}), flush=True)
        del x, y
    for lg in (12, 10, 8, 14, 16, 18, 20, 24): bench(1 << lg, (1 << 28) >> lg)
    print(torch.cuda.get_device_properties(0))
except Exception as e:
    print("cufft yardstick failed:", e)
