"""Latency of the one-shot host API on small transforms (BASELINE config 1 and neighbours): us per fft_auto call.
usage: lat.py [n ...]   (FFTB200_NO_ZEROCOPY=1 for the staged-copy path)"""
import json, os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import numpy as np
import fftb200_loader
F = fftb200_loader.load()
F.require_gpu()
for n in [int(a) for a in sys.argv[1:]] or [64, 1024, 4096, 8192]:
    x = (np.random.default_rng(0).standard_normal(n) + 1j * np.random.default_rng(1).standard_normal(n))
    y = F.fft_auto(x)
    err = float(np.linalg.norm(y - np.fft.fft(x)) / np.linalg.norm(y))
    out = np.empty_like(x)
    L = F.lib
    for _ in range(50): L.fft_auto(F.ptr(x), F.ptr(out), n, -1)
    t0 = time.perf_counter()
    for _ in range(500): L.fft_auto(F.ptr(x), F.ptr(out), n, -1)
    us = (time.perf_counter() - t0) / 500 * 1e6
    print(json.dumps({"n": n, "us_per_fft_auto": round(us, 1), "rel_err_vs_numpy": err, "zero_copy": not os.environ.get("FFTB200_NO_ZEROCOPY")}), flush=True)
