"""Reference-oracle fixture for BASELINE config 4 (one 2^30-point transform, SURVEY.md 8d cfg4: "budget one host oracle run").

Runs the UNMODIFIED reference split_radix_fft (algorithms/core/split_radix.c:58-70, oracle/_ref/libfftref.so) in place on the
2^30-point synthetic input (seed 45, the counter-based stream of SURVEY.md 8d) - 16 GiB, single-threaded, minutes - and keeps
what a GPU box needs to check a distributed result without holding the 16 GiB reference:

  * exact values of a strided sample of bins: every 2^10-th bin (2^20 bins) -> oracle_2p30_strided.npy (16 MiB, git-ignored:
    regenerate with this script), and every 2^14-th bin (2^16 bins) -> oracle_2p30_strided_small.npy (committed);
  * a sketch of the WHOLE vector: for each chunk of 2^20 consecutive bins, K = 8 inner products with pseudo-random +-1 vectors
    (sign of element i under vector j = hash bit of (i, j), tests/sketch.py) and the chunk energy. For a result Y,
    E |sketch_j(Y) - sketch_j(X)|^2 = ||Y - X||^2 over the chunk, so 8192 numbers estimate the full-vector relative L2 error
    to ~1.6 % (relative standard deviation sqrt(2 / 8192) of the squared error) -> oracle_2p30_sketch.npz (committed).

usage: python tests/golden/make_oracle_2p30.py [log_n]   (log_n < 30 writes oracle_2p<log_n>_*.np* for the CPU self-test)
"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle import oracle as O   # noqa: E402
import sketch                     # noqa: E402

SEED = 45


def main():
    log_n = int(sys.argv[1]) if len(sys.argv) > 1 else 30
    n = 1 << log_n
    here = os.path.dirname(os.path.abspath(__file__))
    p, r = O.port(), O.ref()
    t0 = time.time()
    x = p.fill(SEED, 0, n)
    print(f"input 2^{log_n} points generated in {time.time() - t0:.1f} s", flush=True)
    t0 = time.time()
    r.lib.split_radix_fft(O._ptr(x.view(np.float64)), n, -1)   # in place
    print(f"reference split_radix_fft in {time.time() - t0:.1f} s", flush=True)
    tag = f"oracle_2p{log_n}"
    step_big, step_small = 1 << max(0, log_n - 20), 1 << max(0, log_n - 16)
    if log_n >= 30:   # 16 MiB, git-ignored; the smaller sizes keep the 1 MiB sample only
        np.save(os.path.join(here, tag + "_strided.npy"), np.ascontiguousarray(x[::step_big]))
    np.save(os.path.join(here, tag + "_strided_small.npy"), np.ascontiguousarray(x[::step_small]))
    t0 = time.time()
    sk, en = sketch.sketch_numpy(x, first=0)
    np.savez(os.path.join(here, tag + "_sketch.npz"), sketch=sk, energy=en, log_n=log_n, seed=SEED,
             log_chunk=sketch.LOG_CHUNK, k=sketch.K)
    print(f"sketch of {sk.shape[0]} chunks in {time.time() - t0:.1f} s; total energy {en.sum():.6e}", flush=True)


if __name__ == "__main__":
    main()
