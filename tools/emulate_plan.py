"""numpy emulation of the tile-kernel index algebra (development aid, CPU only).

Mirrors csrc/fft_tile.cuh: flat tiles in layout idx = c + (N'/M)*k, DIT sub-passes with the
reference's per-stage twiddle tables (flat offset Mt*h + kappa - 1), F / MID / LAST global passes.
Checked against the oracle so that the CUDA kernel only has to reproduce these formulas.
"""
import sys, os
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
from oracle import oracle as O


def bitrev(x, bits):
    r = 0
    for _ in range(bits):
        r = (r << 1) | (x & 1)
        x >>= 1
    return r


def subpass(state, Np, C, R, M_loc, M_glob, kappa0, tab, trivial):
    """One in-CTA sub-pass over a flat tile of Np elements (C = contiguous columns)."""
    r = R.bit_length() - 1
    S = Np // (M_loc * R)
    Mt = M_glob * M_loc
    out = np.empty_like(state)
    for u in range(Np // R):
        cp, kloc = u % S, u // S
        col = cp % C
        kappa = kappa0(col) + M_glob * kloc
        v = [state[cp + rho * S + (Np // M_loc) * kloc] for rho in range(R)]
        w = [v[bitrev(b, r)] for b in range(R)]
        for s in range(1, r + 1):
            hh = 1 << (s - 1)
            for base in range(0, R, 2 * hh):
                for q in range(hh):
                    h = hh + q
                    if trivial:
                        tw = np.exp(-2j * np.pi * q / (2 * hh))
                    else:
                        tw = tab[Mt * h + kappa - 1]
                    t = w[base + q + hh] * tw
                    w[base + q + hh] = w[base + q] - t
                    w[base + q] = w[base + q] + t
        for q in range(R):
            out[u + (Np // R) * q] = w[q]
    return out


def run_tile(flat, C, radices, M_glob, kappa0, tab, first_trivial):
    Np = flat.size
    M_loc = 1
    for i, R in enumerate(radices):
        flat = subpass(flat, Np, C, R, M_loc, M_glob, kappa0, tab, first_trivial and i == 0)
        M_loc *= R
    return flat


def fft_plan(x, passes, C, tab):
    """passes: list of radix lists, e.g. [[4,4],[4,4]] for N=256 as 16 x 16."""
    N = x.size
    cur = x.copy()
    M = 1
    for pi, radices in enumerate(passes):
        P = int(np.prod(radices))
        nxt = np.empty_like(cur)
        rest = N // (M * P)  # number of c' values
        if rest == 1 and len(passes) > 1:
            # LAST: tile over k
            for k0 in range(0, M, C):
                flat = np.empty(P * C, complex)
                for col in range(C):
                    for rho in range(P):
                        flat[rho * C + col] = cur[rho + P * (k0 + col)]
                flat = run_tile(flat, C, radices, M, lambda col: k0 + col, tab, False)
                for col in range(C):
                    for q in range(P):
                        nxt[(k0 + col) + M * q] = flat[q * C + col]
        else:
            Ct = min(C, rest)
            for k in range(M):
                for c0 in range(0, rest, Ct):
                    flat = np.empty(P * Ct, complex)
                    for col in range(Ct):
                        for rho in range(P):
                            flat[rho * Ct + col] = cur[(c0 + col) + rho * rest + (N // M) * k]
                    flat = run_tile(flat, Ct, radices, M, lambda col: k, tab, M == 1)
                    for col in range(Ct):
                        for q in range(P):
                            nxt[(c0 + col) + rest * (k + M * q)] = flat[q * Ct + col]
        cur = nxt
        M *= P
    return cur


if __name__ == "__main__":
    p = O.port()
    for N, passes, C in [(64, [[8, 8]], 1), (256, [[16, 16]], 1), (128, [[2, 8, 8]], 1),
                         (256, [[4, 4], [4, 4]], 4), (512, [[8], [8], [8]], 4),
                         (1024, [[4, 4], [2, 4], [8]], 4), (2048, [[4, 8], [8, 8]], 8)]:
        x = p.fill(7, 0, N)
        tab = p.twiddle_tables(N)
        y = fft_plan(x, passes, C, tab)
        print(N, passes, C, "vs oracle", O.rel_l2(y, p.fft(x)), "vs numpy", O.rel_l2(y, np.fft.fft(x)))
