"""Shared-memory bank-conflict check for the tile kernel's padded layout (development aid).
For 16-byte accesses a warp is served in quarter-warps of 8 lanes; a quarter is conflict-free when
its 8 element indices are distinct mod 8 (8 x 16 B = all 32 banks)."""
import itertools, sys

def degree(idxs):
    worst = 0
    for q in range(0, 32, 8):
        lanes = idxs[q:q + 8]
        cnt = {}
        for i in set(lanes):
            cnt[i % 8] = cnt.get(i % 8, 0) + 1
        worst = max(worst, max(cnt.values()))
    return worst

def analyse(logp, logc, loge, nt, mode, radices, psh):
    lognp = logp + logc
    NP, E = 1 << lognp, 1 << loge
    T = NP // E
    pos = lambda i: i + (i >> psh)
    res = []
    lm = 0
    for I, lr in enumerate(radices):
        R = 1 << lr; NB = E // R
        LS = lognp - lm - lr; S = 1 << LS
        worst_r = worst_w = 1
        for w0 in range(0, min(T, 64), 32):
            lanes = list(range(w0, min(w0 + 32, T)))
            if len(lanes) < 32:   # several sub-tiles share a warp: same pattern, offset by SM_STRIDE
                lanes = [l % T for l in range(32)]
                offs = [(l // T) * (NP + (NP >> psh) + 2) for l in range(32)]
            else:
                offs = [0] * 32
            for b in range(NB):
                us = []
                for t in lanes:
                    tau = t + T * b
                    if I == 0 and mode == 'last':
                        rowlow, col = tau & ((1 << (logp - lr)) - 1), tau >> (logp - lr)
                        us.append(col + (rowlow << logc))
                    else:
                        us.append(tau)
                if I > 0:
                    for rho in range(R):
                        idx = [pos((u & (S - 1)) + ((u >> LS) << (lognp - lm)) + rho * S) + o for u, o in zip(us, offs)]
                        worst_r = max(worst_r, degree(idx))
                if I < len(radices) - 1:
                    for q in range(R):
                        idx = [pos(u + q * (NP >> lr)) + o for u, o in zip(us, offs)]
                        worst_w = max(worst_w, degree(idx))
        res.append((worst_r, worst_w))
        lm += lr
    return res

CONFIGS = [
    # logp, logc, loge, nt, mode, radices
    (5, 0, 3, 32, 'contig', [2, 3]), (6, 0, 3, 16, 'contig', [3, 3]), (7, 0, 4, 16, 'contig', [3, 4]),
    (8, 0, 4, 16, 'contig', [4, 4]), (9, 0, 3, 4, 'contig', [3, 3, 3]), (10, 0, 4, 4, 'contig', [2, 4, 4]),
    (11, 0, 4, 2, 'contig', [3, 4, 4]), (12, 0, 4, 1, 'contig', [4, 4, 4]), (13, 0, 4, 1, 'contig', [3, 3, 3, 4]),
    (6, 4, 4, 1, 'strided', [3, 3]), (7, 4, 4, 1, 'strided', [3, 4]), (8, 4, 4, 1, 'strided', [4, 4]),
    (9, 3, 4, 1, 'strided', [3, 3, 3]),
    (6, 4, 4, 1, 'last', [3, 3]), (7, 4, 4, 1, 'last', [3, 4]), (8, 4, 4, 1, 'last', [4, 4]),
    (9, 3, 4, 1, 'last', [3, 3, 3]),
]
if __name__ == '__main__':
    for cfg in CONFIGS:
        line = []
        for psh in (2, 3, 4, 5, 6, 31):
            r = analyse(*cfg, psh)
            line.append("psh%d:%s" % (psh, "/".join("%d,%d" % x for x in r)))
        print(cfg, " ".join(line))
