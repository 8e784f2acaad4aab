"""Swizzle search for the fused two-pass kernel (fft_fused.cuh) - development aid.

A tile is 4096 complex doubles, logical index I = lo + 2^LB * f + 2^(LB+LP) * hi, where f is the LP-bit
transform field (Stockham state after `a` stages: f = c'' + 2^(LP-a) * k), lo / hi are batch bits
(pass A: lo = column, no hi; pass B: no lo, hi = k-column). Sub-pass (a, r): butterfly (lo, c'', kloc, hi)
gathers f = c'' + 2^(LP-a-r) * rho + 2^(LP-a) * kloc and scatters f = c'' + 2^(LP-a-r) * (kloc + 2^a * q).
Thread t does butterflies u = t + 256 * bb; u's bits are (ORDER_LOW) [lo][c''][kloc][hi] or
(ORDER_HIGH) [hi][lo][c''][kloc], least significant first.
Physical position = I ^ (bit(I,b0) | bit(I,b1) << 1 | bit(I,b2) << 2). A 16-byte access is served per
quarter-warp (8 lanes): conflict-free when the 8 positions are distinct mod 8.
Prints, per configuration and exchange, a conflict-free (b0, b1, b2) (or 'id')."""
import itertools, sys

def fields(u, LB, LP, LH, a, r, order):
    ncpp = LP - a - r
    if order == 'low':
        lo = u & ((1 << LB) - 1); u >>= LB
        cpp = u & ((1 << ncpp) - 1); u >>= ncpp
        kloc = u & ((1 << a) - 1); u >>= a
        hi = u
    else:
        hi = u & ((1 << LH) - 1); u >>= LH
        lo = u & ((1 << LB) - 1); u >>= LB
        cpp = u & ((1 << ncpp) - 1); u >>= ncpp
        kloc = u
    return lo, cpp, kloc, hi

def gidx(u, rho, LB, LP, LH, a, r, order):
    lo, cpp, kloc, hi = fields(u, LB, LP, LH, a, r, order)
    return lo + ((cpp + (rho << (LP - a - r)) + (kloc << (LP - a))) << LB) + (hi << (LB + LP))

def sidx(u, q, LB, LP, LH, a, r, order):
    lo, cpp, kloc, hi = fields(u, LB, LP, LH, a, r, order)
    return lo + ((cpp + ((kloc + (q << a)) << (LP - a - r))) << LB) + (hi << (LB + LP))

def swz(I, b):
    if b is None: return I
    return I ^ (((I >> b[0]) & 1) | (((I >> b[1]) & 1) << 1) | (((I >> b[2]) & 1) << 2))

def degree(pos):
    worst = 1
    for w in range(0, len(pos), 8):
        cnt = {}
        for p in set(pos[w:w + 8]): cnt[p % 8] = cnt.get(p % 8, 0) + 1
        worst = max(worst, max(cnt.values()))
    return worst

def pattern_sets(cfg):
    """returns per exchange e (0 = TMA layout before sub-pass 0) the list of lane-index lists"""
    LB, LP, LH, subs = cfg
    out = []
    for j, (a, r, order) in enumerate(subs):
        NB = 16 >> r
        g, s = [], []
        for bb in range(NB):
            for x in range(1 << r):
                g.append([gidx(t + 256 * bb, x, LB, LP, LH, a, r, order) for t in range(256)])
                s.append([sidx(t + 256 * bb, x, LB, LP, LH, a, r, order) for t in range(256)])
        out.append((g, s))
    return out

def search(cfg, verbose=True):
    pats = pattern_sets(cfg)
    nsub = len(pats)
    res = []
    cands = [None] + [(s, s + 1, s + 2) for s in range(3, 10)] + [c for c in itertools.permutations(range(3, 12), 3)]
    for e in range(nsub):          # layout read by sub-pass e (written by sub-pass e-1, or the TMA for e = 0)
        reads = pats[e][0]
        writes = pats[e - 1][1] if e > 0 else []
        best = None
        for c in ([None] if e == 0 else cands):
            d = max([degree([swz(i, c) for i in lanes]) for lanes in reads + writes])
            if best is None or d < best[0]: best = (d, c)
            if d == 1: break
        res.append(best)
    return res

def subs_for(LP, kind):
    # radix split: r0 = LP - 8 for LP >= 9 (then 4, 4); LP <= 8: (LP - 4, 4)
    rs = [LP - 8, 4, 4] if LP >= 9 else [LP - 4, 4]
    subs, a = [], 0
    for j, r in enumerate(rs):
        order = 'high' if (kind == 'B' and j == len(rs) - 1) else 'low'
        subs.append((a, r, order)); a += r
    return subs

if __name__ == '__main__':
    for kind in 'AB':
        for LP in range(6, 11):
            LB, LH = (12 - LP, 0) if kind == 'A' else (0, 12 - LP)
            cfg = (LB, LP, LH, subs_for(LP, kind))
            r = search(cfg)
            print(kind, 'LP=%d' % LP, 'subs', [(a, rr, o) for a, rr, o in cfg[3]], '->',
                  ' '.join('e%d:%s(deg%d)' % (e, 'id' if c is None else '%d,%d,%d' % c, d) for e, (d, c) in enumerate(r)))
