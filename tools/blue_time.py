"""Device-timed Bluestein plans (development): blue_time.py [n] [batches...]; honours FFTB200_NO_FUSED_CHIRP for A/B."""
import ctypes as C, json, os, sys
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import fftb200_loader
F = fftb200_loader.load(); L = F.lib
F.require_gpu()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1000003
for b in [int(a) for a in sys.argv[2:]] or [1, 16]:
    p = F.engine_plan(n, b, F.FFTB200_BLUESTEIN)
    din = L.fftb200_malloc(16 * n * b); dout = L.fftb200_malloc(16 * n * b)
    L.fftb200_fill_splitmix(din, 46, 0, n * b)
    for _ in range(3): L.fftb200_plan_exec(p, din, dout)
    ts = []; ms = C.c_float()
    for _ in range(20):
        L.fftb200_timer_start(p); L.fftb200_plan_exec_async(p, din, dout); L.fftb200_timer_stop(p, C.byref(ms)); ts.append(ms.value)
    ts.sort()
    print(json.dumps({"n": n, "batch": b, "fused_chirp": not os.environ.get("FFTB200_NO_FUSED_CHIRP"), "ms_best": round(ts[0], 4), "ms_med": round(ts[10], 4),
                      "launches": L.fftb200_plan_launches(p), "plan": L.fftb200_plan_describe(p).decode()}), flush=True)
    L.fftb200_plan_destroy(p); L.fftb200_free(din); L.fftb200_free(dout)
