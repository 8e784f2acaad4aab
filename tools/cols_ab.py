import ctypes as C, json, os, sys
sys.path.insert(0, "/root/repo")
import fftb200_loader
F = fftb200_loader.load(); L = F.lib
F.require_gpu()
for lg in (22, 23, 24, 25):
  for batch in (1, 2):
    n = 1 << lg; tot = n * batch
    m_in = L.fft_gpu_alloc(tot); m_out = L.fft_gpu_alloc(tot)
    L.fftb200_fill_splitmix(L.fftb200_devptr_of(m_in), 43, 0, tot)
    res = {}
    for tag, env in (("cols", {}), ("three_pass", {"FFTB200_NO_FUSED_COLS": "1"})):
        os.environ.pop("FFTB200_NO_FUSED_COLS", None); os.environ.update(env)
        plan = L.fft_gpu_plan_1d(n, batch, -1); eng = L.fftb200_engine_of(plan)
        din, dout = L.fftb200_devptr_of(m_in), L.fftb200_devptr_of(m_out)
        ts = []; ms = C.c_float()
        for i in range(15):
            L.fftb200_timer_start(eng); L.fftb200_plan_exec_async(eng, din, dout); L.fftb200_timer_stop(eng, C.byref(ms))
            if i >= 3: ts.append(ms.value)
        res[tag] = round(min(ts), 4)
        L.fft_gpu_destroy_plan(plan)
    print(json.dumps({"log_n": lg, "batch": batch, **res}), flush=True)
    L.fft_gpu_free(m_in); L.fft_gpu_free(m_out)
