"""ctypes binding of lib/libfft_b200.so - the B200 drop-in for the reference's FFT hot path.

The product is the shared library (C99 host API + sm_100a kernels); this module only loads it and
mirrors the reference's public C API (include/fft_auto.h, include/fft_gpu.h) one to one, so tests and
bench.py call exactly what a C program linked against the library would call. There is no Python or
CPU implementation behind these names: if the library is missing, import fails.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("FFTB200_LIB") or os.path.join(_HERE, "lib", "libfft_b200.so")  # env: A/B builds

FFT_FORWARD, FFT_INVERSE = -1, 1
FFT_GPU_CUDA, FFT_GPU_AUTO = 1, -1
FFT_ESTIMATE, FFT_PREFER_GPU = 0, 1 << 9
FFTB200_C2C, FFTB200_BLUESTEIN, FFTB200_R2C, FFTB200_C2R = 0, 1, 2, 3

_vp, _dp = C.c_void_p, C.POINTER(C.c_double)

if not os.path.exists(LIB_PATH):
    raise ImportError("libfft_b200.so is not built: run `python fft-implementation-in-c_b200/build.py` "
                      "(there is no fallback implementation)")
lib = C.CDLL(LIB_PATH)

_SIGS = {
    # include/fft_gpu.h
    "fft_gpu_init": (C.c_int, [C.c_int]),
    "fft_gpu_cleanup": (None, []),
    "fft_gpu_available": (C.c_int, []),
    "fft_gpu_get_backend": (C.c_int, []),
    "fft_gpu_alloc": (_vp, [C.c_size_t]),
    "fft_gpu_free": (None, [_vp]),
    "fft_gpu_copy_h2d": (None, [_vp, _vp, C.c_size_t]),
    "fft_gpu_copy_d2h": (None, [_vp, _vp, C.c_size_t]),
    "fft_gpu_plan_1d": (_vp, [C.c_int, C.c_int, C.c_int]),
    "fft_gpu_execute": (None, [_vp, _vp, _vp]),
    "fft_gpu_destroy_plan": (None, [_vp]),
    "fft_gpu_dft_1d": (C.c_int, [_vp, _vp, C.c_int, C.c_int]),
    "fft_gpu_dft_1d_batch": (C.c_int, [_vp, _vp, C.c_int, C.c_int, C.c_int]),
    "fft_gpu_plan_2d": (_vp, [C.c_int, C.c_int, C.c_int]),
    "fft_gpu_dft_2d": (C.c_int, [_vp, _vp, C.c_int, C.c_int, C.c_int]),
    "fft_gpu_get_device_name": (C.c_char_p, []),
    "fft_gpu_get_memory_info": (None, [C.POINTER(C.c_size_t), C.POINTER(C.c_size_t)]),
    "fft_gpu_set_device": (C.c_int, [C.c_int]),
    # include/fft_auto.h
    "fft_plan_dft_1d": (_vp, [C.c_int, _vp, _vp, C.c_int, C.c_uint]),
    "fft_execute": (None, [_vp]),
    "fft_execute_dft": (None, [_vp, _vp, _vp]),
    "fft_destroy_plan": (None, [_vp]),
    "fft_auto": (C.c_int, [_vp, _vp, C.c_int, C.c_int]),
    "fft_plan_r2c_1d": (_vp, [C.c_int, _vp, _vp, C.c_uint]),
    "fft_plan_c2r_1d": (_vp, [C.c_int, _vp, _vp, C.c_uint]),
    "fft_plan_dft_2d": (_vp, [C.c_int, C.c_int, _vp, _vp, C.c_int, C.c_uint]),
    "fft_export_wisdom_to_string": (_vp, []),
    "fft_import_wisdom_from_string": (C.c_int, [C.c_char_p]),
    "fft_get_hardware_capabilities": (C.c_uint, []),
    "fft_plan_with_nthreads": (None, [C.c_int]),
    "fft_alloc_complex": (_vp, [C.c_size_t]),
    "fft_alloc_real": (_vp, [C.c_size_t]),
    "fft_free": (None, [_vp]),
    "fft_version": (C.c_char_p, []),
    # include/fftb200.h (engine C-ABI)
    "fftb200_device_count": (C.c_int, []),
    "fftb200_set_device": (C.c_int, [C.c_int]),
    "fftb200_get_device": (C.c_int, []),
    "fftb200_device_name": (C.c_char_p, []),
    "fftb200_mem_info": (C.c_int, [C.POINTER(C.c_size_t), C.POINTER(C.c_size_t)]),
    "fftb200_sm_count": (C.c_int, []),
    "fftb200_device_reset": (C.c_int, []),
    "fftb200_malloc": (_vp, [C.c_size_t]),
    "fftb200_free": (None, [_vp]),
    "fftb200_host_alloc": (_vp, [C.c_size_t]),
    "fftb200_host_free": (None, [_vp]),
    "fftb200_memcpy_h2d": (C.c_int, [_vp, _vp, C.c_size_t]),
    "fftb200_memcpy_d2h": (C.c_int, [_vp, _vp, C.c_size_t]),
    "fftb200_memcpy_d2d": (C.c_int, [_vp, _vp, C.c_size_t]),
    "fftb200_memset": (C.c_int, [_vp, C.c_int, C.c_size_t]),
    "fftb200_fill_splitmix": (C.c_int, [_vp, C.c_ulonglong, C.c_ulonglong, C.c_ulonglong]),
    "fftb200_plan_create": (C.c_int, [C.POINTER(_vp), _vp]),
    "fftb200_plan_exec": (C.c_int, [_vp, _vp, _vp]),
    "fftb200_plan_exec_async": (C.c_int, [_vp, _vp, _vp]),
    "fftb200_plan_sync": (C.c_int, [_vp]),
    "fftb200_plan_exec_host": (C.c_int, [_vp, _vp, _vp]),
    "fftb200_plan_destroy": (None, [_vp]),
    "fftb200_plan_launches": (C.c_int, [_vp]),
    "fftb200_plan_describe": (C.c_char_p, [_vp]),
    "fftb200_timer_start": (C.c_int, [_vp]),
    "fftb200_timer_stop": (C.c_int, [_vp, C.POINTER(C.c_float)]),
    "fftb200_pointwise_mul": (C.c_int, [_vp, _vp, _vp, C.c_size_t]),
    "fftb200_pointwise_mul_conj": (C.c_int, [_vp, _vp, _vp, C.c_size_t]),
    "fftb200_plan_set_stream": (C.c_int, [_vp, _vp]),
    "fftb200_transpose": (C.c_int, [_vp, _vp, C.c_longlong, C.c_longlong, C.c_longlong, _vp]),
    "fftb200_plan_create_partial": (C.c_int, [C.POINTER(_vp), _vp, C.c_int, C.c_int, C.c_int, C.c_double]),
    "fftb200_permute_bac": (C.c_int, [_vp, _vp, C.c_longlong, C.c_longlong, C.c_longlong, _vp]),
    "fftb200_plan_stream": (_vp, [_vp]),
    "fftb200_peers_create": (C.c_int, [C.POINTER(_vp), C.POINTER(_vp), C.c_int, C.c_int]),
    "fftb200_peers_destroy": (None, [_vp]),
    "fftb200_plan_set_peer_output": (C.c_int, [_vp, _vp, C.c_int, C.c_int]),
    "fftb200_push_columns": (C.c_int, [_vp, _vp, _vp, C.c_longlong, C.c_int]),
    "fftb200_ipc_export": (C.c_int, [_vp, _vp]),
    "fftb200_ipc_open": (_vp, [_vp]),
    "fftb200_ipc_close": (C.c_int, [_vp]),
    "fftb200_last_error": (C.c_char_p, []),
    # host-library helpers (not part of the reference API)
    "fftb200_engine_of": (_vp, [_vp]),
    "fftb200_devptr_of": (_vp, [_vp]),
    "fftb200_host_twiddles": (_vp, [C.c_int]),
    "fftb200_host_twiddles_accurate": (_vp, [C.POINTER(C.c_int)]),
    "fftb200_host_chirp": (None, [_vp, C.c_int, C.c_int]),
    "fftb200_host_tables_release": (None, []),
    "fftb200_host_twiddles_dist": (C.c_int, [_vp, C.c_int, C.c_int, C.c_int, C.c_int]),
    "fftb200_shard_range": (C.c_int, [C.c_longlong, C.c_int, C.c_int, C.POINTER(C.c_longlong), C.POINTER(C.c_longlong)]),
    "fftb200_host_set_gpus": (None, [C.c_int]),
    "fftb200_host_get_gpus": (C.c_int, []),
    "fftb200_host_cache_stats": (None, [C.POINTER(C.c_longlong), C.POINTER(C.c_longlong)]),
    "fftb200_barrier_create": (C.c_int, [C.POINTER(_vp), C.POINTER(_vp), C.c_int, C.c_int]),
    "fftb200_barrier_enqueue": (C.c_int, [_vp, _vp]),
    "fftb200_barrier_destroy": (None, [_vp]),
    "fftb200_stream_sync": (C.c_int, [_vp]),
    "fftb200_enable_peer_access": (C.c_int, [C.c_int]),
    # include/fftb200_dist.h (argtypes with the callback are set by dist.py)
    "fftb200_dist_choose_split": (C.c_int, [C.c_int, C.c_int]),
    "fftb200_dist_sync": (C.c_int, [_vp]),
    "fftb200_dist_stream": (_vp, [_vp]),
    "fftb200_dist_log_m": (C.c_int, [_vp]),
    "fftb200_dist_describe": (C.c_char_p, [_vp]),
    "fftb200_dist_destroy": (None, [_vp]),
    "fft_gpu_convolution": (C.c_int, [_vp, C.c_int, _vp, C.c_int, _vp]),
    "fft_gpu_circular_convolution": (C.c_int, [_vp, _vp, C.c_int, _vp]),
    "fft_gpu_cross_correlation": (C.c_int, [_vp, _vp, C.c_int, _vp]),
    "fft_gpu_autocorrelation": (C.c_int, [_vp, C.c_int, _vp]),
}
EXPORTS = sorted(_SIGS)
for _name, (_res, _args) in _SIGS.items():
    try:
        _f = getattr(lib, _name)  # AttributeError here = the library does not export a declared symbol
    except AttributeError:
        if os.environ.get("FFTB200_LIB"):   # an A/B build of an older commit (tools/ab_build.sh): time what it has
            continue
        raise
    _f.restype, _f.argtypes = _res, _args


def ptr(a):
    """Address of a contiguous numpy array as void*."""
    assert a.flags["C_CONTIGUOUS"]
    return a.ctypes.data


def require_gpu():
    if lib.fft_gpu_available() != 1:
        raise RuntimeError("no CUDA device: the B200 FFT path has no CPU fallback")
    if lib.fft_gpu_init(FFT_GPU_AUTO) != 0:
        raise RuntimeError("fft_gpu_init failed: " + lib.fftb200_last_error().decode())


def fft_auto(x, sign=-1):
    """fft_auto(in, out, n, sign) on a 1-D complex128 array; returns the output array."""
    x = np.ascontiguousarray(x, dtype=np.complex128)
    out = np.empty_like(x)
    if lib.fft_auto(ptr(x), ptr(out), x.size, sign) != 0:
        raise RuntimeError("fft_auto failed: " + lib.fftb200_last_error().decode())
    return out


def gpu_fft_batch(x, direction=FFT_FORWARD, inplace=False):
    """fft_gpu_alloc / copy_h2d / plan_1d(n, batch) / execute / copy_d2h on a 2-D complex128 array."""
    x = np.ascontiguousarray(x, dtype=np.complex128)
    batch, n = x.shape
    require_gpu()
    m_in = lib.fft_gpu_alloc(x.size)
    m_out = m_in if inplace else lib.fft_gpu_alloc(x.size)
    plan = lib.fft_gpu_plan_1d(n, batch, direction)
    if not m_in or not m_out or not plan:
        raise RuntimeError("fft_gpu setup failed: " + lib.fftb200_last_error().decode())
    try:
        lib.fft_gpu_copy_h2d(m_in, ptr(x), x.size)
        lib.fft_gpu_execute(plan, m_in, m_out)
        out = np.empty_like(x)
        lib.fft_gpu_copy_d2h(ptr(out), m_out, x.size)
    finally:
        lib.fft_gpu_destroy_plan(plan)
        lib.fft_gpu_free(m_in)
        if not inplace:
            lib.fft_gpu_free(m_out)
    return out


def r2c(x):
    """fft_plan_r2c_1d + fft_execute on a 1-D float64 array; returns n/2+1 bins."""
    x = np.ascontiguousarray(x, dtype=np.float64)
    out = np.empty(x.size // 2 + 1, dtype=np.complex128)
    plan = lib.fft_plan_r2c_1d(x.size, ptr(x), ptr(out), 0)
    if not plan:
        raise RuntimeError("fft_plan_r2c_1d failed: " + lib.fftb200_last_error().decode())
    lib.fft_execute(plan)
    lib.fft_destroy_plan(plan)
    return out


def c2r(half, n):
    """fft_plan_c2r_1d + fft_execute: n/2+1 complex bins -> n reals (inverse of r2c, scaled 1/n)."""
    half = np.ascontiguousarray(half, dtype=np.complex128)
    assert half.size == n // 2 + 1
    out = np.empty(n, dtype=np.float64)
    plan = lib.fft_plan_c2r_1d(n, ptr(half), ptr(out), 0)
    if not plan:
        raise RuntimeError("fft_plan_c2r_1d failed: " + lib.fftb200_last_error().decode())
    lib.fft_execute(plan)
    lib.fft_destroy_plan(plan)
    return out


def fft2d(x, sign=-1, api="plan"):
    """2-D transform of a (rows, cols) complex128 array through fft_plan_dft_2d + fft_execute (api="plan"),
    fft_gpu_dft_2d (api="dft") or the device handle API fft_gpu_plan_2d / fft_gpu_execute (api="gpu")."""
    x = np.ascontiguousarray(x, dtype=np.complex128)
    rows, cols = x.shape
    out = np.empty_like(x)
    if api == "plan":
        plan = lib.fft_plan_dft_2d(rows, cols, ptr(x), ptr(out), sign, 0)
        if not plan:
            raise RuntimeError("fft_plan_dft_2d failed: " + lib.fftb200_last_error().decode())
        lib.fft_execute(plan)
        lib.fft_destroy_plan(plan)
    elif api == "dft":
        if lib.fft_gpu_dft_2d(ptr(x), ptr(out), rows, cols, -1 if sign < 0 else 1) != 0:
            raise RuntimeError("fft_gpu_dft_2d failed: " + lib.fftb200_last_error().decode())
    else:
        require_gpu()
        plan = lib.fft_gpu_plan_2d(rows, cols, -1 if sign < 0 else 1)
        mem = lib.fft_gpu_alloc(x.size)
        if not plan or not mem:
            raise RuntimeError("fft_gpu_plan_2d failed: " + lib.fftb200_last_error().decode())
        lib.fft_gpu_copy_h2d(mem, ptr(x), x.size)
        lib.fft_gpu_execute(plan, mem, mem)
        lib.fft_gpu_copy_d2h(ptr(out), mem, x.size)
        lib.fft_gpu_destroy_plan(plan)
        lib.fft_gpu_free(mem)
    return out


def _pair(fn, x, y, n_out):
    x = np.ascontiguousarray(x, dtype=np.complex128)
    y = np.ascontiguousarray(y, dtype=np.complex128)
    out = np.empty(n_out, dtype=np.complex128)
    return x, y, out


def convolution(x, h):
    """fft_gpu_convolution: linear convolution, len(x) + len(h) - 1 samples."""
    x, h, y = _pair(None, x, h, len(x) + len(h) - 1)
    if lib.fft_gpu_convolution(ptr(x), x.size, ptr(h), h.size, ptr(y)) != 0:
        raise RuntimeError("fft_gpu_convolution failed: " + lib.fftb200_last_error().decode())
    return y


def circular_convolution(x, h):
    x, h, y = _pair(None, x, h, len(x))
    if lib.fft_gpu_circular_convolution(ptr(x), ptr(h), x.size, ptr(y)) != 0:
        raise RuntimeError("fft_gpu_circular_convolution failed: " + lib.fftb200_last_error().decode())
    return y


def cross_correlation(x, y):
    x, y, r = _pair(None, x, y, len(x))
    if lib.fft_gpu_cross_correlation(ptr(x), ptr(y), x.size, ptr(r)) != 0:
        raise RuntimeError("fft_gpu_cross_correlation failed: " + lib.fftb200_last_error().decode())
    return r


def autocorrelation(x):
    x = np.ascontiguousarray(x, dtype=np.complex128)
    r = np.empty_like(x)
    if lib.fft_gpu_autocorrelation(ptr(x), x.size, ptr(r)) != 0:
        raise RuntimeError("fft_gpu_autocorrelation failed: " + lib.fftb200_last_error().decode())
    return r


class PlanDesc(C.Structure):
    """struct fftb200_plan_desc (include/fftb200.h)"""
    _fields_ = [("n", C.c_int), ("batch", C.c_int), ("direction", C.c_int), ("kind", C.c_int),
                ("twiddles", C.c_void_p), ("table_n", C.c_int), ("chirp", C.c_void_p),
                ("twiddles_accurate", C.c_void_p), ("accurate_n", C.c_int), ("flags", C.c_uint)]


def engine_plan(n, batch, kind=FFTB200_C2C, direction=FFT_FORWARD):
    """fftb200_plan_create with the host tables the C99 host library would pass (host/fft_gpu.c): returns the plan handle.
    R2C / BLUESTEIN have no batched entry point in the reference's public API; the engine plan is the batched form."""
    m = n
    keep = None
    d = PlanDesc(n, batch, direction, kind, None, 0, None, None, 0, 0)
    if kind == FFTB200_BLUESTEIN:
        m = 1
        while m < 2 * n - 1:
            m <<= 1
        keep = host_chirp(n, direction)
        d.chirp = keep.ctypes.data
    d.twiddles, d.table_n = lib.fftb200_host_twiddles(m), m
    an = C.c_int()
    d.twiddles_accurate = lib.fftb200_host_twiddles_accurate(C.byref(an))
    d.accurate_n = an.value
    plan = _vp()
    if lib.fftb200_plan_create(C.byref(plan), C.byref(d)) != 0:
        raise RuntimeError("fftb200_plan_create failed: " + lib.fftb200_last_error().decode())
    return plan


def host_twiddles(n):
    """The host-built forward stage tables for power-of-two n (n-1 complex), as a numpy copy."""
    p = lib.fftb200_host_twiddles(n)
    if not p:
        raise RuntimeError("fftb200_host_twiddles failed")
    buf = (C.c_double * (2 * (n - 1))).from_address(p)
    return np.frombuffer(buf, dtype=np.complex128).copy()


def host_chirp(n, direction=-1):
    c = np.empty(n, dtype=np.complex128)
    lib.fftb200_host_chirp(ptr(c), n, direction)
    return c


def shard_range(batch, world, rank):
    """(first, count) of the contiguous batch range owned by `rank` (fftb200_shard_range)."""
    first, count = C.c_longlong(), C.c_longlong()
    if lib.fftb200_shard_range(batch, world, rank, C.byref(first), C.byref(count)) != 0:
        raise ValueError("bad shard arguments")
    return first.value, count.value
