"""Mismatch of the Hermitian r2c schedule against the oracle, per size (development tool behind R2C_HERM_MAX_LOG in csrc/fft_plan.cu).
usage: python tools/r2c_herm_table.py  -> markdown rows: log2 n | rel L2 vs oracle (Hermitian) | (full pass B) | ms Hermitian | ms full"""
import ctypes as C, os, sys
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import numpy as np, torch
import fftb200_loader
from oracle import oracle as O
F = fftb200_loader.load(); L = F.lib
F.require_gpu()
p = O.port()
print("| log2 n | batch | rel L2 vs oracle, Hermitian schedule | rel L2 vs oracle, full pass B | ms Hermitian (2^28 reals) | ms full |")
print("|---|---|---|---|---|---|")
for lg in range(14, 21):
    n = 1 << lg; batch = (1 << 28) >> lg
    rows = [0, batch - 1]
    xs = {r: p.fill(47, r * n // 2, n // 2).view(np.float64).copy() for r in rows}
    want = np.stack([p.r2c(xs[r]) for r in rows])
    xd = torch.empty(batch * n, dtype=torch.float64, device="cuda")
    L.fftb200_fill_splitmix(xd.data_ptr(), 47, 0, batch * n // 2)
    yd = torch.zeros(batch, n // 2 + 1, dtype=torch.complex128, device="cuda")
    res = []
    for herm in ("1", "0"):
        os.environ["FFTB200_R2C_HERMITIAN"] = herm
        plan = F.engine_plan(n, batch, F.FFTB200_R2C)
        ms = C.c_float(); ts = []
        for i in range(8):
            L.fftb200_timer_start(plan); assert L.fftb200_plan_exec_async(plan, xd.data_ptr(), yd.data_ptr()) == 0; L.fftb200_timer_stop(plan, C.byref(ms))
            if i >= 3: ts.append(ms.value)
        got = yd[rows].cpu().numpy()
        res.append((O.rel_l2(got, want), min(ts)))
        L.fftb200_plan_destroy(plan)
    print("| %d | %d | %.2e | %.2e | %.3f | %.3f |" % (lg, batch, res[0][0], res[1][0], res[0][1], res[1][1]), flush=True)
