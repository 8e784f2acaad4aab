"""Development check (gpurun): fused r2c plans (N = 2^13 .. 2^20): parity vs the oracle (host path and device batches), timing."""
import ctypes as C, json, math, os, sys
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import numpy as np, torch
import fftb200_loader
from oracle import oracle as O
F = fftb200_loader.load(); L = F.lib
F.require_gpu()
p = O.port()
sys.path.insert(0, fftb200_loader.PKG_DIR)
import importlib.util
spec = importlib.util.spec_from_file_location("fft_b200_dist", os.path.join(fftb200_loader.PKG_DIR, "dist.py"))
D = importlib.util.module_from_spec(spec); spec.loader.exec_module(D)
L.fftb200_host_twiddles_accurate.restype = C.c_void_p
def r2c_plan(n, batch):
    d = D.PlanDesc(n, batch, -1, F.FFTB200_R2C, L.fftb200_host_twiddles(n), n, None, None, 0, 0)
    an = C.c_int(); d.twiddles_accurate = L.fftb200_host_twiddles_accurate(C.byref(an)); d.accurate_n = an.value
    pl = C.c_void_p()
    assert L.fftb200_plan_create(C.byref(pl), C.byref(d)) == 0, L.fftb200_last_error()
    return pl
for lg in [int(a) for a in sys.argv[1:]]:
    n = 1 << lg
    xr = p.fill(47, 0, n).real.copy()
    print(json.dumps({"n": n, "host_path_err": float(O.rel_l2(F.r2c(xr), p.r2c(xr)))}), flush=True)
    for b in (3, 41):
        if n * b > (1 << 24): continue
        x = p.fill(47, 0, n * b).real.copy().reshape(b, n)
        pl = r2c_plan(n, b)
        xd = torch.from_numpy(x).cuda(); yd = torch.zeros((b, n // 2 + 1), dtype=torch.complex128, device="cuda")
        assert L.fftb200_plan_exec(pl, xd.data_ptr(), yd.data_ptr()) == 0
        want = np.stack([p.r2c(r) for r in x])
        print(json.dumps({"n": n, "b": b, "err": float(O.rel_l2(yd.cpu().numpy(), want)), "plan": L.fftb200_plan_describe(pl).decode()[:90]}), flush=True)
        L.fftb200_plan_destroy(pl)
    b = (1 << 28) >> lg
    for env in (None, "1"):
        if env: os.environ["FFTB200_NO_FUSED_R2C"] = "1"
        else: os.environ.pop("FFTB200_NO_FUSED_R2C", None)
        pl = r2c_plan(n, b)
        xd = torch.rand((b, n), dtype=torch.float64, device="cuda"); yd = torch.empty((b, n // 2 + 1), dtype=torch.complex128, device="cuda")
        for _ in range(3): L.fftb200_plan_exec(pl, xd.data_ptr(), yd.data_ptr())
        ts = []; ms = C.c_float()
        for _ in range(6):
            L.fftb200_timer_start(pl); L.fftb200_plan_exec_async(pl, xd.data_ptr(), yd.data_ptr()); L.fftb200_timer_stop(pl, C.byref(ms)); ts.append(ms.value)
        t = min(ts)
        print(json.dumps({"n": n, "batch": b, "fused_r2c": env is None, "ms": round(t, 3), "algorithmic_GBps": round((8 * n + 16 * (n // 2 + 1)) * b / t * 1e-6)}), flush=True)
        L.fftb200_plan_destroy(pl); del xd, yd
    os.environ.pop("FFTB200_NO_FUSED_R2C", None)
