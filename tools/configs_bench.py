"""All BASELINE.json configurations that fit one GPU, device-timed through the engine C-ABI (development / profiles).
cfg1 N=1024 single (fft_auto end to end + device only), cfg2 N=4096 x 65536, the 2^10..2^20 target band at 2^28 points,
cfg3 N=2^24 single and batch 16, cfg5 Bluestein 1000003 (batch 1, 16) and r2c 2^20 x 256."""
import ctypes as C, json, math, os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import numpy as np
import fftb200_loader
sys.path.insert(0, fftb200_loader.PKG_DIR)
F = fftb200_loader.load(); L = F.lib
F.require_gpu()
import importlib.util
spec = importlib.util.spec_from_file_location("fft_b200_dist", os.path.join(fftb200_loader.PKG_DIR, "dist.py"))
D = importlib.util.module_from_spec(spec); spec.loader.exec_module(D)

def time_engine(eng, din, dout, reps=10, warm=3):
    for _ in range(warm): L.fftb200_plan_exec(eng, din, dout)
    ts = []; ms = C.c_float()
    for _ in range(reps):
        L.fftb200_timer_start(eng); L.fftb200_plan_exec_async(eng, din, dout); L.fftb200_timer_stop(eng, C.byref(ms)); ts.append(ms.value)
    return min(ts), sorted(ts)[len(ts) // 2]

def c2c(n, batch, tag):
    tot = n * batch
    m_in = L.fft_gpu_alloc(tot); m_out = L.fft_gpu_alloc(tot)
    L.fftb200_fill_splitmix(L.fftb200_devptr_of(m_in), 43, 0, tot)
    plan = L.fft_gpu_plan_1d(n, batch, -1)
    eng = L.fftb200_engine_of(plan)
    best, med = time_engine(eng, L.fftb200_devptr_of(m_in), L.fftb200_devptr_of(m_out))
    desc = L.fftb200_plan_describe(eng).decode()
    L.fft_gpu_destroy_plan(plan); L.fft_gpu_free(m_in); L.fft_gpu_free(m_out)
    lg = math.log2(n) if n & (n - 1) == 0 else math.log2(n)
    print(json.dumps({"cfg": tag, "n": n, "batch": batch, "ms_best": round(best, 4), "ms_med": round(med, 4), "gflops": round(5 * n * lg * batch / best * 1e-6),
                      "strict_GBps": round(32 * n * batch / best * 1e-6), "plan": desc}), flush=True)

def engine_plan(n, batch, kind):
    d = D.PlanDesc(n, batch, -1, kind, None, 0, None, None, 0, 0)
    m = n
    if kind == F.FFTB200_BLUESTEIN:
        m = 1
        while m < 2 * n - 1: m <<= 1
        chirp = F.host_chirp(n, -1)
        d.chirp = chirp.ctypes.data
    d.twiddles = L.fftb200_host_twiddles(m); d.table_n = m
    acc_n = C.c_int()
    L.fftb200_host_twiddles_accurate.restype = C.c_void_p
    d.twiddles_accurate = L.fftb200_host_twiddles_accurate(C.byref(acc_n)); d.accurate_n = acc_n.value
    p = C.c_void_p()
    assert L.fftb200_plan_create(C.byref(p), C.byref(d)) == 0, L.fftb200_last_error()
    return p

# cfg1
x = np.random.default_rng(0).standard_normal(1024) + 0j
F.fft_auto(x)
t0 = time.perf_counter()
for _ in range(200): F.fft_auto(x)
print(json.dumps({"cfg": "1: fft_auto(1024) host pointers end to end (plan + H2D + kernel + D2H)", "us_per_call": round((time.perf_counter() - t0) / 200 * 1e6, 1)}), flush=True)
c2c(1024, 1, "1: N=1024 single, device only")
c2c(4096, 65536, "2: N=4096 x 65536")
for lg in range(10, 21): c2c(1 << lg, (1 << 28) >> lg, "band 2^10..2^20 at 2^28 points")
c2c(1 << 24, 1, "3: N=2^24 single"); c2c(1 << 24, 16, "3: N=2^24 x 16")
# cfg5: Bluestein
for b in (1, 16):
    n = 1000003
    p = engine_plan(n, b, F.FFTB200_BLUESTEIN)
    din = L.fftb200_malloc(16 * n * b); dout = L.fftb200_malloc(16 * n * b)
    L.fftb200_fill_splitmix(din, 46, 0, n * b)
    best, med = time_engine(p, din, dout)
    print(json.dumps({"cfg": "5: Bluestein n=1000003", "batch": b, "ms_best": round(best, 4), "ms_per_transform": round(best / b, 4), "strict_GBps(32n)": round(32 * n * b / best * 1e-6),
                      "plan": L.fftb200_plan_describe(p).decode()}), flush=True)
    L.fftb200_plan_destroy(p); L.fftb200_free(din); L.fftb200_free(dout)
# cfg5: r2c
n, b = 1 << 20, 256
p = engine_plan(n, b, F.FFTB200_R2C)
din = L.fftb200_malloc(8 * n * b); dout = L.fftb200_malloc(16 * (n // 2 + 1) * b)
L.fftb200_fill_splitmix(din, 47, 0, n * b // 2)
best, med = time_engine(p, din, dout)
print(json.dumps({"cfg": "5: r2c n=2^20 x 256", "ms_best": round(best, 4), "algorithmic_GBps(8n+16(n/2+1))": round((8 * n + 16 * (n // 2 + 1)) * b / best * 1e-6),
                  "gflops(2.5 n log2 n)": round(2.5 * n * 20 * b / best * 1e-6), "plan": L.fftb200_plan_describe(p).decode()}), flush=True)
L.fftb200_plan_destroy(p)
L.fftb200_free(din); L.fftb200_free(dout)
# the real transforms either side of config 5: r2c / c2r at 4096 (inside the pipe kernel) and c2r at 2^20
for lg, kinds in ((12, (F.FFTB200_R2C, F.FFTB200_C2R)), (20, (F.FFTB200_C2R,))):
    n = 1 << lg; b = (1 << 28) >> lg; nh = n // 2 + 1
    dre = L.fftb200_malloc(8 * n * b); dcx = L.fftb200_malloc(16 * nh * b)
    L.fftb200_fill_splitmix(dre, 47, 0, n * b // 2); L.fftb200_fill_splitmix(dcx, 48, 0, nh * b)
    for kind in kinds:
        p = F.engine_plan(n, b, kind, -1 if kind == F.FFTB200_R2C else 1)
        din, dout = (dre, dcx) if kind == F.FFTB200_R2C else (dcx, dre)
        best, med = time_engine(p, din, dout)
        print(json.dumps({"cfg": "%s n=2^%d x %d" % ("r2c" if kind == F.FFTB200_R2C else "c2r", lg, b), "ms_best": round(best, 4),
                          "algorithmic_GBps(8n+16(n/2+1))": round((8 * n + 16 * nh) * b / best * 1e-6), "plan": L.fftb200_plan_describe(p).decode()}), flush=True)
        L.fftb200_plan_destroy(p)
    L.fftb200_free(dre); L.fftb200_free(dcx)
