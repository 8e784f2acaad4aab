"""Device-timed batched r2c / c2r engine plans (development). usage: real_time.py [log_n] [batch]"""
import ctypes as C, json, os, sys
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import fftb200_loader
F = fftb200_loader.load(); L = F.lib
F.require_gpu()
lg = int(sys.argv[1]) if len(sys.argv) > 1 else 20
n = 1 << lg
b = int(sys.argv[2]) if len(sys.argv) > 2 else (1 << 28) >> lg
for kind, name in ((F.FFTB200_R2C, "r2c"), (F.FFTB200_C2R, "c2r")):
    p = F.engine_plan(n, b, kind, -1 if kind == F.FFTB200_R2C else 1)
    nh = n // 2 + 1
    dre = L.fftb200_malloc(8 * n * b); dcx = L.fftb200_malloc(16 * nh * b)
    L.fftb200_fill_splitmix(dre, 47, 0, n * b // 2)
    L.fftb200_fill_splitmix(dcx, 48, 0, nh * b)
    din, dout = (dre, dcx) if kind == F.FFTB200_R2C else (dcx, dre)
    for _ in range(3): L.fftb200_plan_exec(p, din, dout)
    ts = []; ms = C.c_float()
    for _ in range(10):
        L.fftb200_timer_start(p); L.fftb200_plan_exec_async(p, din, dout); L.fftb200_timer_stop(p, C.byref(ms)); ts.append(ms.value)
    ts.sort()
    print(json.dumps({"kind": name, "n": n, "batch": b, "ms_best": round(ts[0], 4), "algorithmic_GBps(8n+16(n/2+1))": round((8 * n + 16 * nh) * b / ts[0] * 1e-6),
                      "plan": L.fftb200_plan_describe(p).decode()}), flush=True)
    L.fftb200_plan_destroy(p); L.fftb200_free(dre); L.fftb200_free(dcx)
