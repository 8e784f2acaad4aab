"""In-library multi-device fan-out (fftb200_host_set_gpus / FFTB200_GPUS) measured end to end from ONE process: host buffers from
fft_alloc_complex, fft_gpu_dft_1d_batch(in, out, n, batch, FFT_FORWARD) with H2D + kernels + D2H inside the timed region.
usage: python tools/fanout_e2e.py [n] [batch]   (development tool; the table goes to profiles/r02_multigpu.md)"""
import ctypes as C, json, os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import numpy as np
import fftb200_loader
from oracle import oracle as O
F = fftb200_loader.load(); L = F.lib
F.require_gpu()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
batch = int(sys.argv[2]) if len(sys.argv) > 2 else 2 * 65536
total = n * batch
hin, hout = L.fft_alloc_complex(total), L.fft_alloc_complex(total)
assert hin and hout
x = np.ctypeslib.as_array(C.cast(hin, C.POINTER(C.c_double)), shape=(2 * total,))
rng = np.random.default_rng(1)
blk = rng.standard_normal(1 << 22)
for i in range(0, 2 * total, 1 << 22):
    x[i:i + (1 << 22)] = blk[: min(1 << 22, 2 * total - i)]
ndev = L.fftb200_device_count()
p = O.port()
for g in [g for g in (1, 2, 4, 8) if g <= ndev]:
    L.fftb200_host_set_gpus(g)
    assert L.fft_gpu_dft_1d_batch(hin, hout, n, batch, -1) == 0   # builds the per-device plans and staging rings
    ts = []
    for _ in range(3):
        t0 = time.perf_counter()
        assert L.fft_gpu_dft_1d_batch(hin, hout, n, batch, -1) == 0
        ts.append(time.perf_counter() - t0)
    y = np.ctypeslib.as_array(C.cast(hout, C.POINTER(C.c_double)), shape=(2 * total,)).view(np.complex128)
    rows = [0, batch // 2 + 1, batch - 1]
    err = max(O.rel_l2(y[r * n:(r + 1) * n], p.fft(x.view(np.complex128)[r * n:(r + 1) * n], -1)) for r in rows)
    dt = min(ts)
    print(json.dumps({"gpus": g, "n": n, "batch": batch, "ms": round(dt * 1e3, 2), "gflops": round(5 * n * np.log2(n) * batch / dt * 1e-9, 1),
                      "GBps_each_way": round(16 * total / dt * 1e-9, 1), "rel_l2_vs_oracle": err}), flush=True)
L.fftb200_host_set_gpus(1)
L.fft_free(hin); L.fft_free(hout)
