"""ctypes loaders for the CPU oracle (TEST INFRASTRUCTURE ONLY).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import this.
The product package never does.

  port()  -> oracle/liboracle.so       our C restatement (oracle/fft_oracle.c)
  ref()   -> oracle/_ref/libfftref.so  the UNMODIFIED reference CPU library (built by oracle/Makefile
             from /root/reference; prebuilt artefact on the GPU box)
  par()   -> oracle/_ref/libparref.so  the reference's pthreads/OpenMP path (CPU baseline timing only)
  simd()  -> oracle/_ref/libsimdref.so the reference's SSE2 float path (CPU baseline timing only; numerically wrong)
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_dp = C.POINTER(C.c_double)


def build(quiet=True):
    """Compile liboracle.so and, when /root/reference is present, oracle/_ref/*.so."""
    subprocess.run(["make", "-C", _HERE, "all"], check=True,
                   stdout=subprocess.DEVNULL if quiet else None)


def _ptr(a):
    assert a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(_dp)


class Port:
    """Our restatement (oracle/fft_oracle.c)."""

    def __init__(self):
        path = os.path.join(_HERE, "liboracle.so")
        if not os.path.exists(path):
            build()
        self.lib = L = C.CDLL(path)
        L.oracle_fft_pow2.argtypes = [_dp, C.c_int, C.c_int, C.c_int]
        L.oracle_fft_pow2_batch.argtypes = [_dp, C.c_int, C.c_long, C.c_int, C.c_int]
        L.oracle_twiddle_tables.argtypes = [_dp, C.c_int]
        L.oracle_chirp.argtypes = [_dp, C.c_int, C.c_int]
        L.oracle_bluestein.argtypes = [_dp, C.c_int, C.c_int]
        L.oracle_fft_auto.argtypes = [_dp, _dp, C.c_int, C.c_int, C.c_int]
        L.oracle_r2c.argtypes = [_dp, _dp, C.c_int]
        L.oracle_naive_dft.argtypes = [_dp, _dp, C.c_int, C.c_int]
        L.oracle_fill.argtypes = [_dp, C.c_uint64, C.c_uint64, C.c_uint64]
        L.oracle_fill.restype = None
        L.oracle_c2r.argtypes = [_dp, _dp, C.c_int]
        L.oracle_fft2d.argtypes = [_dp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]
        L.oracle_convolution.argtypes = [_dp, C.c_int, _dp, C.c_int, _dp]
        L.oracle_circular_convolution.argtypes = [_dp, _dp, C.c_int, _dp]
        L.oracle_cross_correlation.argtypes = [_dp, _dp, C.c_int, _dp]
        L.oracle_autocorrelation.argtypes = [_dp, C.c_int, _dp]

    def fft(self, x, sign=-1, quirk=False):
        """fft_auto semantics on a 1-D complex128 array (any n): returns a new array."""
        x = np.ascontiguousarray(x, dtype=np.complex128)
        out = np.empty_like(x)
        rc = self.lib.oracle_fft_auto(_ptr(x.view(np.float64)), _ptr(out.view(np.float64)),
                                      x.size, sign, int(quirk))
        if rc != 0:
            raise ValueError("oracle_fft_auto failed")
        return out

    def fft_batch(self, x, sign=-1, threads=1):
        """Batched pow2 c2c over the last axis of a 2-D complex128 array (in a copy)."""
        x = np.array(x, dtype=np.complex128, order="C", copy=True)
        batch, n = x.shape
        rc = self.lib.oracle_fft_pow2_batch(_ptr(x.view(np.float64)), n, batch,
                                            -1 if sign < 0 else 1, threads)
        if rc != 0:
            raise ValueError("oracle_fft_pow2_batch: n must be a power of two")
        return x

    def fft_batch_inplace(self, x, sign=-1, threads=1):
        batch, n = x.shape
        return self.lib.oracle_fft_pow2_batch(_ptr(x.view(np.float64)), n, batch,
                                              -1 if sign < 0 else 1, threads)

    def twiddle_tables(self, n):
        t = np.empty(n - 1, dtype=np.complex128)
        self.lib.oracle_twiddle_tables(_ptr(t.view(np.float64)), n)
        return t

    def chirp(self, n, direction=-1):
        c = np.empty(n, dtype=np.complex128)
        self.lib.oracle_chirp(_ptr(c.view(np.float64)), n, direction)
        return c

    def r2c(self, x):
        x = np.ascontiguousarray(x, dtype=np.float64)
        out = np.empty(x.size // 2 + 1, dtype=np.complex128)
        if self.lib.oracle_r2c(_ptr(x), _ptr(out.view(np.float64)), x.size) != 0:
            raise ValueError("oracle_r2c failed")
        return out

    def c2r(self, half, n):
        half = np.ascontiguousarray(half, dtype=np.complex128)
        assert half.size == n // 2 + 1
        out = np.empty(n, dtype=np.float64)
        if self.lib.oracle_c2r(_ptr(half.view(np.float64)), _ptr(out), n) != 0:
            raise ValueError("oracle_c2r failed")
        return out

    def fft2d(self, x, sign=-1, double_scale=False, quirk=False):
        """Row-column 2-D transform of a (rows, cols) complex128 array (in a copy)."""
        x = np.array(x, dtype=np.complex128, order="C", copy=True)
        rows, cols = x.shape
        if self.lib.oracle_fft2d(_ptr(x.view(np.float64)), rows, cols, -1 if sign < 0 else 1, int(double_scale), int(quirk)) != 0:
            raise ValueError("oracle_fft2d: rows and cols must be powers of two")
        return x

    def convolution(self, x, h):
        x = np.ascontiguousarray(x, dtype=np.complex128); h = np.ascontiguousarray(h, dtype=np.complex128)
        y = np.empty(x.size + h.size - 1, dtype=np.complex128)
        if self.lib.oracle_convolution(_ptr(x.view(np.float64)), x.size, _ptr(h.view(np.float64)), h.size, _ptr(y.view(np.float64))) != 0:
            raise ValueError("oracle_convolution failed")
        return y

    def circular_convolution(self, x, h):
        x = np.ascontiguousarray(x, dtype=np.complex128); h = np.ascontiguousarray(h, dtype=np.complex128)
        y = np.empty_like(x)
        if self.lib.oracle_circular_convolution(_ptr(x.view(np.float64)), _ptr(h.view(np.float64)), x.size, _ptr(y.view(np.float64))) != 0:
            raise ValueError("oracle_circular_convolution failed")
        return y

    def cross_correlation(self, x, y):
        x = np.ascontiguousarray(x, dtype=np.complex128); y = np.ascontiguousarray(y, dtype=np.complex128)
        r = np.empty_like(x)
        if self.lib.oracle_cross_correlation(_ptr(x.view(np.float64)), _ptr(y.view(np.float64)), x.size, _ptr(r.view(np.float64))) != 0:
            raise ValueError("oracle_cross_correlation failed")
        return r

    def autocorrelation(self, x):
        return self.cross_correlation(x, x)

    def naive_dft(self, x, sign=-1):
        x = np.ascontiguousarray(x, dtype=np.complex128)
        out = np.empty_like(x)
        self.lib.oracle_naive_dft(_ptr(x.view(np.float64)), _ptr(out.view(np.float64)), x.size,
                                  -1 if sign < 0 else 1)
        return out

    def fill(self, seed, first, count):
        """Synthetic input stream: `count` complex elements starting at element `first`."""
        x = np.empty(count, dtype=np.complex128)
        self.lib.oracle_fill(_ptr(x.view(np.float64)), seed, first, count)
        return x


class Ref:
    """The unmodified reference library (oracle/_ref/libfftref.so)."""

    def __init__(self):
        path = os.path.join(_HERE, "_ref", "libfftref.so")
        if not os.path.exists(path):
            raise FileNotFoundError(path)
        self.lib = L = C.CDLL(path)
        L.fft_auto.argtypes = [_dp, _dp, C.c_int, C.c_int]
        for name in ("radix2_dit_fft", "radix4_fft", "split_radix_fft", "bluestein_fft"):
            getattr(L, name).argtypes = [_dp, C.c_int, C.c_int]
            getattr(L, name).restype = None
        L.naive_dft.argtypes = [_dp, _dp, C.c_int, C.c_int]
        L.naive_dft.restype = None

    def fft_auto(self, x, sign=-1):
        x = np.ascontiguousarray(x, dtype=np.complex128)
        out = np.empty_like(x)
        if self.lib.fft_auto(_ptr(x.view(np.float64)), _ptr(out.view(np.float64)), x.size, sign) != 0:
            raise ValueError("reference fft_auto failed")
        return out

    def inplace(self, name, x, direction=-1):
        x = np.array(x, dtype=np.complex128, order="C", copy=True)
        getattr(self.lib, name)(_ptr(x.view(np.float64)), x.size, direction)
        return x


class Par:
    """Reference multi-threaded CPU path (optimizations/parallel_fft.c), CPU-baseline timing only."""

    def __init__(self):
        path = os.path.join(_HERE, "_ref", "libparref.so")
        if not os.path.exists(path):
            raise FileNotFoundError(path)
        self.lib = L = C.CDLL(path)
        for name in ("oracle_ref_four_step", "oracle_ref_radix2_parallel"):
            getattr(L, name).argtypes = [_dp, C.c_int, C.c_int, C.c_int]
            getattr(L, name).restype = None
        L.oracle_ref_batch_execute.argtypes = [_dp, _dp, C.c_int, C.c_long, C.c_int, C.c_int]

    def batch_execute(self, x, out, sign, threads):
        """Reference fft_plan_dft_1d / fft_execute_dft over a (batch, n) array, one plan per thread."""
        batch, n = x.shape
        return self.lib.oracle_ref_batch_execute(_ptr(x.view(np.float64)), _ptr(out.view(np.float64)), n, batch,
                                                 sign, threads)

    def radix2_parallel(self, x, direction, threads):
        self.lib.oracle_ref_radix2_parallel(_ptr(x.view(np.float64)), x.size, direction, threads)

    def four_step(self, x, direction, threads):
        self.lib.oracle_ref_four_step(_ptr(x.view(np.float64)), x.size, direction, threads)


class Simd:
    """Reference SIMD path (optimizations/simd_fft.c: fft_radix2_sse2, float, numerically wrong - SURVEY 6.2): TIMING ONLY."""

    def __init__(self):
        path = os.path.join(_HERE, "_ref", "libsimdref.so")
        if not os.path.exists(path):
            raise FileNotFoundError(path)
        self.lib = L = C.CDLL(path)
        L.oracle_ref_sse2_time.argtypes = [C.c_int, C.c_int]
        L.oracle_ref_sse2_time.restype = C.c_double

    def sse2_seconds_per_transform(self, n, reps):
        return float(self.lib.oracle_ref_sse2_time(n, reps))


class Apps:
    """The reference's FFT callers (applications/convolution.c, image_fft.c, power_spectrum.c), unmodified."""

    def __init__(self):
        path = os.path.join(_HERE, "_ref", "libappsref.so")
        if not os.path.exists(path):
            raise FileNotFoundError(path)
        self.lib = L = C.CDLL(path)
        L.oracle_ref_fft_convolution.argtypes = [_dp, C.c_int, _dp, C.c_int, _dp]
        L.oracle_ref_circular_convolution.argtypes = [_dp, _dp, C.c_int, _dp]
        L.oracle_ref_fft_2d.argtypes = [_dp, C.c_int, C.c_int, C.c_int]
        L.oracle_ref_autocorrelation.argtypes = [_dp, C.c_int, _dp]
        L.oracle_ref_cross_correlation.argtypes = [_dp, _dp, C.c_int, _dp]
        for f in ("oracle_ref_fft_convolution", "oracle_ref_circular_convolution", "oracle_ref_fft_2d",
                  "oracle_ref_autocorrelation", "oracle_ref_cross_correlation"):
            getattr(L, f).restype = None

    def convolution(self, x, h):
        x = np.ascontiguousarray(x, dtype=np.complex128); h = np.ascontiguousarray(h, dtype=np.complex128)
        y = np.empty(x.size + h.size - 1, dtype=np.complex128)
        self.lib.oracle_ref_fft_convolution(_ptr(x.view(np.float64)), x.size, _ptr(h.view(np.float64)), h.size, _ptr(y.view(np.float64)))
        return y

    def circular_convolution(self, x, h):
        x = np.ascontiguousarray(x, dtype=np.complex128); h = np.ascontiguousarray(h, dtype=np.complex128)
        y = np.empty_like(x)
        self.lib.oracle_ref_circular_convolution(_ptr(x.view(np.float64)), _ptr(h.view(np.float64)), x.size, _ptr(y.view(np.float64)))
        return y

    def fft2d(self, x, sign=-1):
        x = np.array(x, dtype=np.complex128, order="C", copy=True)
        self.lib.oracle_ref_fft_2d(_ptr(x.view(np.float64)), x.shape[0], x.shape[1], -1 if sign < 0 else 1)
        return x

    def cross_correlation(self, x, y):
        x = np.ascontiguousarray(x, dtype=np.complex128); y = np.ascontiguousarray(y, dtype=np.complex128)
        r = np.empty_like(x)
        self.lib.oracle_ref_cross_correlation(_ptr(x.view(np.float64)), _ptr(y.view(np.float64)), x.size, _ptr(r.view(np.float64)))
        return r

    def autocorrelation(self, x):
        x = np.ascontiguousarray(x, dtype=np.complex128)
        r = np.empty_like(x)
        self.lib.oracle_ref_autocorrelation(_ptr(x.view(np.float64)), x.size, _ptr(r.view(np.float64)))
        return r


_cache = {}


def apps():
    if "apps" not in _cache:
        _cache["apps"] = Apps()
    return _cache["apps"]


def port():
    if "port" not in _cache:
        _cache["port"] = Port()
    return _cache["port"]


def ref():
    if "ref" not in _cache:
        _cache["ref"] = Ref()
    return _cache["ref"]


def par():
    if "par" not in _cache:
        _cache["par"] = Par()
    return _cache["par"]


def simd():
    if "simd" not in _cache:
        _cache["simd"] = Simd()
    return _cache["simd"]


def have_ref():
    return os.path.exists(os.path.join(_HERE, "_ref", "libfftref.so"))


def rel_l2(a, b):
    a = np.asarray(a).ravel()
    b = np.asarray(b).ravel()
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300))
