"""Generates tests/golden/reference_outputs.npz from the UNMODIFIED reference library.

Run in the build container (needs oracle/_ref/libfftref.so, which oracle/Makefile compiles from
/root/reference): `python tests/golden/make_golden.py`. Inputs are the counter-based splitmix64 stream
of SURVEY.md 8(d) (oracle_fill: seed, element index), so only outputs are stored. Sizes up to 4096 are
stored whole; larger ones as a strided sample plus the L2 norm of the whole output.

Entry points used: fft_auto (reference algorithms/auto/fft_auto.c:325), which routes power-of-two n to
radix2_dit_fft / radix4_fft / split_radix_fft and everything else relevant here to bluestein_fft.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle as O  # noqa: E402

FULL = [(2, 42), (32, 42), (64, 42), (128, 42), (256, 42), (512, 42), (1024, 42), (2048, 43), (4096, 43),
        (3, 46), (6, 46), (12, 46), (97, 46), (360, 46), (1009, 46), (4099, 46)]
QUIRK = [(4, 42), (8, 42), (16, 42)]           # the reference skips the bit reversal here (fft_common.h:59-77)
SAMPLED = [(1 << 13, 44), (1 << 14, 44), (1 << 16, 44), (1 << 17, 44), (1 << 20, 44), (100003, 46), (1000003, 46)]
STRIDE_SAMPLES = 2048


def main():
    p, r = O.port(), O.ref()
    out = {}
    for n, seed in FULL + QUIRK:
        x = p.fill(seed, 0, n)
        for sign, tag in ((-1, "f"), (1, "i")):
            out[f"full_{n}_{seed}_{tag}"] = r.fft_auto(x, sign)
    for n, seed in SAMPLED:
        x = p.fill(seed, 0, n)
        idx = (np.arange(STRIDE_SAMPLES, dtype=np.int64) * (n // STRIDE_SAMPLES + 1) * 7 + 3) % n
        for sign, tag in ((-1, "f"), (1, "i")):
            y = r.fft_auto(x, sign)
            out[f"samp_{n}_{seed}_{tag}"] = y[idx]
            out[f"norm_{n}_{seed}_{tag}"] = np.array([np.linalg.norm(y)])
        out[f"idx_{n}"] = idx
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "reference_outputs.npz")
    np.savez_compressed(path, **out)
    print(path, os.path.getsize(path), "bytes,", len(out), "arrays")


if __name__ == "__main__":
    main()
