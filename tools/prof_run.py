"""FUSED_PROF build (FFTB200_VARIANT=prof FFTB200_NVCC_EXTRA=-DFUSED_PROF python build.py): per-tile cycle accounts of the fused kernel.
usage: FFTB200_LIB=.../ab_prof.so FFTB200_FUSED_PROF_PRINT=1 python tools/prof_run.py 14 16 20"""
import ctypes as C, os, sys
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import fftb200_loader
F = fftb200_loader.load(); L = F.lib
F.require_gpu()
for lg in [int(a) for a in sys.argv[1:]]:
    n = 1 << lg; batch = (1 << 28) >> lg; tot = n * batch
    m_in = L.fft_gpu_alloc(tot); m_out = L.fft_gpu_alloc(tot)
    L.fftb200_fill_splitmix(L.fftb200_devptr_of(m_in), 43, 0, tot)
    plan = L.fft_gpu_plan_1d(n, batch, -1); eng = L.fftb200_engine_of(plan)
    print("== 2^%d:" % lg, L.fftb200_plan_describe(eng).decode(), flush=True)
    ms = C.c_float()
    for i in range(3):
        L.fftb200_timer_start(eng); L.fftb200_plan_exec_async(eng, L.fftb200_devptr_of(m_in), L.fftb200_devptr_of(m_out)); L.fftb200_timer_stop(eng, C.byref(ms))
        print("   %.4f ms" % ms.value, flush=True)
    L.fft_gpu_destroy_plan(plan); L.fft_gpu_free(m_in); L.fft_gpu_free(m_out)
