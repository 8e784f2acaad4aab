"""Generates tests/golden/reference_apps.npz from the UNMODIFIED reference's FFT callers:
applications/convolution.c (fft_convolution, circular_convolution), applications/image_fft.c (fft_2d) and
applications/power_spectrum.c (autocorrelation_fft, cross_correlation_fft), compiled where they lie into
oracle/_ref/libappsref.so by oracle/Makefile (wrapper: oracle/ref_apps_wrap.c).

Run in the build container: `python tests/golden/make_golden_apps.py`. Inputs are the counter-based splitmix64
stream (oracle_fill: seed, element index), so only outputs are stored.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle as O  # noqa: E402

CONV = [(300, 45, 50), (1024, 1024, 51), (5000, 17, 52)]      # (nx, nh, seed)
CIRC = [(256, 53), (4096, 54)]                                # (n, seed)
CORR = [(100, 55), (1024, 56), (3000, 57)]                    # (n, seed)
IMG = [(64, 128, 58), (256, 64, 59), (8, 32, 60), (1024, 128, 61)]   # (rows, cols, seed)


def main():
    p, a = O.port(), O.apps()
    out = {}
    for nx, nh, seed in CONV:
        out[f"conv_{nx}_{nh}_{seed}"] = a.convolution(p.fill(seed, 0, nx), p.fill(seed, 1 << 20, nh))
    for n, seed in CIRC:
        out[f"circ_{n}_{seed}"] = a.circular_convolution(p.fill(seed, 0, n), p.fill(seed, 1 << 20, n))
    for n, seed in CORR:
        out[f"xcorr_{n}_{seed}"] = a.cross_correlation(p.fill(seed, 0, n), p.fill(seed, 1 << 20, n))
        out[f"acorr_{n}_{seed}"] = a.autocorrelation(p.fill(seed, 0, n))
    for rows, cols, seed in IMG:
        x = p.fill(seed, 0, rows * cols).reshape(rows, cols)
        y = a.fft2d(x, -1)
        yi = a.fft2d(x, 1)        # the reference's inverse: scaled by 1/(rows*cols) twice (image_fft.c:63-71)
        if rows * cols > 8192:    # store a strided sample and the norm
            idx = (np.arange(2048, dtype=np.int64) * 37 + 5) % (rows * cols)
            out[f"img_{rows}_{cols}_{seed}_idx"] = idx
            out[f"img_{rows}_{cols}_{seed}_f"] = y.ravel()[idx]
            out[f"img_{rows}_{cols}_{seed}_i"] = yi.ravel()[idx]
            out[f"img_{rows}_{cols}_{seed}_norm"] = np.array([np.linalg.norm(y), np.linalg.norm(yi)])
        else:
            out[f"img_{rows}_{cols}_{seed}_f"] = y
            out[f"img_{rows}_{cols}_{seed}_i"] = yi
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "reference_apps.npz")
    np.savez_compressed(path, **out)
    print(path, os.path.getsize(path), "bytes,", len(out), "arrays")


if __name__ == "__main__":
    main()
