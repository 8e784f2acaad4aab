"""Development (gpurun): per-pass timing of the fused kernel with the dependency counters ignored."""
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
    for dbg, name in ((0, "fused"), (4, "nowait"), (1, "A only"), (2, "B only")):
        os.environ["FFTB200_FUSED_DEBUG"] = str(dbg)
        t = timeit(1 << lg, (1 << 28) >> lg)
        print(json.dumps({"log_n": lg, "mode": name, "ms": round(t, 3), "GBps_32B": round(32 * (1 << 28) / t * 1e-6)}), flush=True)
