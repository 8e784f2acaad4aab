"""Development sweep (gpurun): fused kernel timing vs lag / slots / group size. usage: fused_sweep.py log_n [log_n ...]"""
import ctypes as C, json, math, os, sys
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import fftb200_loader
F = fftb200_loader.load(); L = F.lib
F.require_gpu()
def timeit(n, batch, reps=6):
    tot = n * batch
    m_in = L.fft_gpu_alloc(tot); m_out = L.fft_gpu_alloc(tot)
    L.fftb200_fill_splitmix(L.fftb200_devptr_of(m_in), 43, 0, tot)
    plan = L.fft_gpu_plan_1d(n, batch, -1)
    eng = L.fftb200_engine_of(plan)
    din, dout = L.fftb200_devptr_of(m_in), L.fftb200_devptr_of(m_out)
    for _ in range(2): L.fftb200_plan_exec(eng, din, dout)
    ts = []; ms = C.c_float()
    for _ in range(reps):
        L.fftb200_timer_start(eng); L.fftb200_plan_exec_async(eng, din, dout); L.fftb200_timer_stop(eng, C.byref(ms)); ts.append(ms.value)
    L.fft_gpu_destroy_plan(plan); L.fft_gpu_free(m_in); L.fft_gpu_free(m_out)
    return min(ts)
for lg in [int(a) for a in sys.argv[1:]]:
    tpt = 1 << (lg - 12)
    for gt in sorted(set([max(1, 16 // tpt), max(1, 32 // tpt), max(1, 64 // tpt)])):
        T = gt * tpt
        for lag_tiles in (150, 300, 450, 600):
            lag = max(1, -(-lag_tiles // T))
            for extra in (1, 2, 4):
                os.environ["FFTB200_FUSED_GT"] = str(gt); os.environ["FFTB200_FUSED_LAG"] = str(lag); os.environ["FFTB200_FUSED_SLOTS"] = str(lag + extra)
                t = timeit(1 << lg, (1 << 28) >> lg)
                print(json.dumps({"log_n": lg, "gt": gt, "T": T, "lag": lag, "slots": lag + extra, "scratch_MB": (lag + extra) * T / 16, "ms": round(t, 3), "strict_GBps": round(32 * (1 << 28) / t * 1e-6)}), flush=True)
