"""Timing of the real transforms through the engine C-ABI (development tool): r2c / c2r at 2^28 real points per execution.
usage: [FFTB200_LIB=...] python tools/real_ab.py 14 16 20"""
import ctypes as C, json, os, sys
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import torch
import fftb200_loader
F = fftb200_loader.load(); L = F.lib
F.require_gpu()
for lg in [int(a) for a in sys.argv[1:]]:
    n = 1 << lg; batch = (1 << 28) >> lg
    xr = torch.rand(batch, n, dtype=torch.float64, device="cuda")
    half = torch.zeros(batch, n // 2 + 1, dtype=torch.complex128, device="cuda")
    for kind, name, src, dst, d in ((F.FFTB200_R2C, "r2c", xr, half, -1), (F.FFTB200_C2R, "c2r", half, xr, 1)):
        plan = F.engine_plan(n, batch, kind, d)
        ts = []; ms = C.c_float()
        for i in range(13):
            L.fftb200_timer_start(plan); assert L.fftb200_plan_exec_async(plan, src.data_ptr(), dst.data_ptr()) == 0; L.fftb200_timer_stop(plan, C.byref(ms))
            if i >= 3: ts.append(ms.value)
        print(json.dumps({"lib": os.path.basename(os.environ.get("FFTB200_LIB", "default")), "kind": name, "log_n": lg, "batch": batch, "ms_best": round(min(ts), 4),
                          "GBps(8n+16(n/2+1))": round((8 * n + 16 * (n // 2 + 1)) * batch / min(ts) * 1e-6)}), flush=True)
        L.fftb200_plan_destroy(plan)
