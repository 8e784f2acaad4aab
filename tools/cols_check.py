"""Development check (gpurun): column-mode fused plans (N = 2^22 .. 2^25): parity vs the oracle, in place, timing vs the 3-pass plan."""
import ctypes as C, json, math, os, sys
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import numpy as np
import fftb200_loader
from oracle import oracle as O
F = fftb200_loader.load(); L = F.lib
F.require_gpu()
p = O.port()
def timeit(n, batch, reps=8):
    tot = n * batch
    m_in = L.fft_gpu_alloc(tot); m_out = L.fft_gpu_alloc(tot)
    L.fftb200_fill_splitmix(L.fftb200_devptr_of(m_in), 43, 0, tot)
    plan = L.fft_gpu_plan_1d(n, batch, -1)
    eng = L.fftb200_engine_of(plan)
    din, dout = L.fftb200_devptr_of(m_in), L.fftb200_devptr_of(m_out)
    for _ in range(3): L.fftb200_plan_exec(eng, din, dout)
    ts = []; ms = C.c_float()
    for _ in range(reps):
        L.fftb200_timer_start(eng); L.fftb200_plan_exec_async(eng, din, dout); L.fftb200_timer_stop(eng, C.byref(ms)); ts.append(ms.value)
    desc = L.fftb200_plan_describe(eng).decode()
    L.fft_gpu_destroy_plan(plan); L.fft_gpu_free(m_in); L.fft_gpu_free(m_out)
    t = min(ts)
    return {"n": n, "batch": batch, "ms_best": round(t, 4), "gflops": round(5*n*math.log2(n)*batch/t*1e-6), "strict_GBps": round(32*n*batch/t*1e-6), "plan": desc}
for lg in [int(a) for a in sys.argv[1:]]:
    n = 1 << lg
    for b in (1, 3):
        if n * b > (1 << 26): continue
        x = p.fill(44, 0, n * b).reshape(b, n)
        res = {"n": n, "b": b}
        for d in (-1, 1):
            y = F.gpu_fft_batch(x, d)
            res["fwd" if d < 0 else "inv"] = float(O.rel_l2(y, p.fft_batch(x, d)))
            if d < 0: res["inplace_same"] = bool(np.array_equal(y, F.gpu_fft_batch(x, d, inplace=True)))
        print(json.dumps(res), flush=True)
    for env in (None, "1"):
        if env: os.environ["FFTB200_NO_FUSED_COLS"] = env
        else: os.environ.pop("FFTB200_NO_FUSED_COLS", None)
        for b in (1, (1 << 28) >> lg):
            print(json.dumps(timeit(n, b)), flush=True)
    os.environ.pop("FFTB200_NO_FUSED_COLS", None)
