"""Device-timed c2c plan of a chosen shape (development): time_nb.py log_n batch [reps]; environment switches apply."""
import ctypes as C, json, os, sys
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import fftb200_loader
F = fftb200_loader.load(); L = F.lib
F.require_gpu()
n = 1 << int(sys.argv[1]); batch = int(sys.argv[2]); reps = int(sys.argv[3]) if len(sys.argv) > 3 else 30
tot = n * batch
m_in = L.fft_gpu_alloc(tot); m_out = L.fft_gpu_alloc(tot)
L.fftb200_fill_splitmix(L.fftb200_devptr_of(m_in), 43, 0, tot)
plan = L.fft_gpu_plan_1d(n, batch, -1); eng = L.fftb200_engine_of(plan)
din, dout = L.fftb200_devptr_of(m_in), L.fftb200_devptr_of(m_out)
for _ in range(3): L.fftb200_plan_exec(eng, din, dout)
ts = []; ms = C.c_float()
for _ in range(reps):
    L.fftb200_timer_start(eng); L.fftb200_plan_exec_async(eng, din, dout); L.fftb200_timer_stop(eng, C.byref(ms)); ts.append(ms.value)
ts.sort()
print(json.dumps({"n": n, "batch": batch, "best": round(ts[0], 4), "med": round(ts[len(ts) // 2], 4), "plan": L.fftb200_plan_describe(eng).decode()[:90]}))
