"""Random-sign sketches of very long complex vectors (TEST INFRASTRUCTURE): lets a GPU box check a 2^30-point result against
the reference oracle without holding the 16 GiB reference output (tests/golden/make_oracle_2p30.py).

For each chunk of 2^LOG_CHUNK consecutive elements and each of K sign vectors s_j (s_j[i] = +-1, bit j of a 64-bit hash of the
GLOBAL element index i), the sketch is sum_i s_j[i] * x[i]. For two vectors X, Y:  E_j |sk_j(Y) - sk_j(X)|^2 = ||Y - X||^2
over the chunk (the signs are pairwise independent), so the mean over j and the sum over chunks estimate the squared L2
distance of the full vectors; dividing by the stored energy gives the relative L2 error.

The same hash is written twice: numpy (uint64) for the host oracle and torch (int64, wrapping multiplies, masked shifts) for
device-resident results; tests/test_sketch.py checks they agree bit for bit.
"""
import numpy as np

LOG_CHUNK = 20
K = 8
_C1, _C2 = 0xBF58476D1CE4E5B9, 0x94D049BB133111EB
_GOLD = 0x9E3779B97F4A7C15


def hash_numpy(idx):
    """splitmix64 finaliser of (idx + GOLD) over a uint64 array."""
    with np.errstate(over="ignore"):
        z = idx.astype(np.uint64) + np.uint64(_GOLD)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(_C1)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(_C2)
        z = z ^ (z >> np.uint64(31))
    return z


def signs_numpy(first, count):
    """[K, count] float64 array of +-1 for global indices first .. first + count - 1."""
    h = hash_numpy(np.arange(first, first + count, dtype=np.uint64))
    out = np.empty((K, count), dtype=np.float64)
    for j in range(K):
        out[j] = 1.0 - 2.0 * ((h >> np.uint64(8 * j + 3)) & np.uint64(1)).astype(np.float64)
    return out


def sketch_numpy(x, first=0):
    """x: complex128 slice of the long vector starting at global index `first` (length a multiple of the chunk, or shorter than
    one chunk). Returns (sketch [chunks, K] complex128, energy [chunks] float64)."""
    x = np.asarray(x)
    chunk = min(1 << LOG_CHUNK, x.size)
    assert x.size % chunk == 0
    nch = x.size // chunk
    sk = np.empty((nch, K), dtype=np.complex128)
    en = np.empty(nch, dtype=np.float64)
    for c in range(nch):
        xc = x[c * chunk:(c + 1) * chunk]
        s = signs_numpy(first + c * chunk, chunk)
        sk[c] = s @ xc.real + 1j * (s @ xc.imag)
        en[c] = float(np.vdot(xc, xc).real)
    return sk, en


def _u64(v):
    """Python int (unsigned 64-bit constant) -> the int64 with the same bits."""
    return v - (1 << 64) if v >= (1 << 63) else v


def hash_torch(idx):
    """Same hash on a torch int64 tensor (any device): wrapping multiplies, logical shifts through masks."""
    import torch  # noqa: F401
    def lsr(z, s):
        return (z >> s) & ((1 << (64 - s)) - 1)
    z = idx + _u64(_GOLD)
    z = (z ^ lsr(z, 30)) * _u64(_C1)
    z = (z ^ lsr(z, 27)) * _u64(_C2)
    return z ^ lsr(z, 31)


def sketch_torch(y, first=0):
    """y: complex128 torch tensor (device-resident block starting at global index `first`). Returns numpy (sketch, energy)."""
    import torch
    chunk = min(1 << LOG_CHUNK, y.numel())
    assert y.numel() % chunk == 0
    nch = y.numel() // chunk
    sk = np.empty((nch, K), dtype=np.complex128)
    en = np.empty(nch, dtype=np.float64)
    yr = torch.view_as_real(y)
    per = max(1, min(nch, (1 << 24) // chunk))    # chunks per step: bounded temporaries
    for c0 in range(0, nch, per):
        c1 = min(nch, c0 + per)
        idx = torch.arange(first + c0 * chunk, first + c1 * chunk, dtype=torch.int64, device=y.device)
        h = hash_torch(idx)
        blk = yr[c0 * chunk:c1 * chunk].reshape(c1 - c0, chunk, 2)
        for j in range(K):
            s = (1.0 - 2.0 * ((h >> (8 * j + 3)) & 1).to(torch.float64)).reshape(c1 - c0, chunk, 1)
            v = (blk * s).sum(dim=1).cpu().numpy()
            sk[c0:c1, j] = v[:, 0] + 1j * v[:, 1]
        en[c0:c1] = (blk * blk).sum(dim=(1, 2)).cpu().numpy()
    return sk, en


def rel_l2_estimate(sk_y, sk_x, energy_x):
    """Estimated relative L2 distance of the full vectors from their sketches."""
    d = np.abs(np.asarray(sk_y) - np.asarray(sk_x)) ** 2
    return float(np.sqrt(d.mean(axis=1).sum() / np.asarray(energy_x).sum()))
