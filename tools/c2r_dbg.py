"""Debug: c2r through the fused kernel, half-spectrum input vs the c2r_expand path. usage: c2r_dbg.py log_n batch"""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import numpy as np, torch
import fftb200_loader
F = fftb200_loader.load(); L = F.lib
F.require_gpu()
lg, batch = int(sys.argv[1]), int(sys.argv[2])
n = 1 << lg
rng = np.random.default_rng(0)
x = rng.standard_normal((batch, n))
half = np.fft.rfft(x, axis=1)
hd = torch.from_numpy(half).cuda()
res = {}
for herm in ("0", "1"):
    os.environ["FFTB200_C2R_HERMITIAN"] = herm
    plan = F.engine_plan(n, batch, F.FFTB200_C2R, 1)
    print(herm, L.fftb200_plan_describe(plan).decode(), flush=True)
    yd = torch.full((batch, n), float("nan"), dtype=torch.float64, device="cuda")
    rc = L.fftb200_plan_exec(plan, hd.data_ptr(), yd.data_ptr())
    print("rc", rc, L.fftb200_last_error().decode(), flush=True)
    if rc != 0: sys.exit(1)
    res[herm] = yd.cpu().numpy()
    L.fftb200_plan_destroy(plan)
d = res["1"] - res["0"]
print("max abs diff", np.abs(d).max(), "nan", np.isnan(res["1"]).sum(), "vs x", np.abs(res["0"] - x).max())
bad = np.argwhere(np.abs(d) > 0)
print("mismatches", len(bad), bad[:10].tolist())
