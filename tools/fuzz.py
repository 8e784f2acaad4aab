"""Randomised shape sweep on a B200 (run under gpurun): every kind of plan at random (n, batch, direction, in-place)
against numpy's accurate transform. The reference recurrence is up to ~1e-11 away from the accurate DFT at 2^20 (1.6e-10
at 2^24), so the bound here is loose (gross errors only: wrong tile, race, ragged batch); exact parity is tests/.
usage: fuzz.py [seconds] [seed]"""
import json, os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import numpy as np
import fftb200_loader

F = fftb200_loader.load(); L = F.lib
F.require_gpu()
budget = float(sys.argv[1]) if len(sys.argv) > 1 else 60.0
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 1)


def rel(a, b):
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300))


def tol_for(n):
    return 1e-12 if n <= (1 << 14) else 2e-11 if n <= (1 << 20) else 1e-9


t0 = time.time(); runs = 0; bad = []; worst = {}
while time.time() - t0 < budget:
    kind = rng.choice(["pow2", "pow2", "pow2", "blue", "r2c", "c2r", "2d"])
    try:
        if kind == "pow2":
            lg = int(rng.integers(1, 23))
            n = 1 << lg
            cap = max(1, (1 << 24) >> lg)
            batch = int(min(cap, rng.choice([1, 2, 3, rng.integers(1, 40), rng.integers(1, 700), rng.integers(1, 6000)])))
            d = int(rng.choice([-1, 1])); inplace = bool(rng.integers(0, 2))
            x = (rng.standard_normal((batch, n)) + 1j * rng.standard_normal((batch, n)))
            y = F.gpu_fft_batch(x, d, inplace=inplace)
            want = np.fft.fft(x, axis=1) if d < 0 else np.fft.ifft(x, axis=1)
            e = rel(y, want); tag = ("pow2", lg, batch, d, inplace)
            lim = tol_for(n)
        elif kind == "blue":
            n = int(rng.integers(3, 70000))
            if n & (n - 1) == 0: n += 1
            batch = int(rng.integers(1, 5)); d = int(rng.choice([-1, 1]))
            x = (rng.standard_normal((batch, n)) + 1j * rng.standard_normal((batch, n)))
            y = F.gpu_fft_batch(x, d)
            want = np.fft.fft(x, axis=1) if d < 0 else np.fft.ifft(x, axis=1)
            e = rel(y, want); tag = ("blue", n, batch, d); lim = 1e-10   # the reference recurrence at m = 2^17: ~2e-11
        elif kind == "r2c":
            n = 1 << int(rng.integers(1, 21))
            x = rng.standard_normal(n)
            e = rel(F.r2c(x), np.fft.rfft(x)); tag = ("r2c", n); lim = tol_for(n)
        elif kind == "c2r":
            n = 1 << int(rng.integers(1, 21))
            h = np.fft.rfft(rng.standard_normal(n))
            e = rel(F.c2r(h, n), np.fft.irfft(h, n)); tag = ("c2r", n); lim = tol_for(n)
        else:
            r, c = 1 << int(rng.integers(1, 11)), 1 << int(rng.integers(1, 11))
            if rng.integers(0, 3) == 0: r = int(rng.integers(2, 200))
            d = int(rng.choice([-1, 1]))
            x = (rng.standard_normal((r, c)) + 1j * rng.standard_normal((r, c)))
            y = F.fft2d(x, d, api=str(rng.choice(["plan", "dft", "gpu"])))
            want = np.fft.fft2(x) if d < 0 else np.fft.ifft2(x)
            e = rel(y, want); tag = ("2d", r, c, d); lim = 1e-11
    except Exception as ex:   # noqa: BLE001 - report and keep sweeping
        bad.append({"tag": [str(t) for t in tag] if "tag" in dir() else kind, "error": str(ex)[:200]})
        continue
    runs += 1
    if not (e <= lim):
        bad.append({"tag": [str(t) for t in tag], "err": e, "lim": lim})
    k = tag[0]
    if e > worst.get(k, (0, None))[0]: worst[k] = (e, [str(t) for t in tag])
print(json.dumps({"runs": runs, "seconds": round(time.time() - t0, 1), "bad": bad[:20], "n_bad": len(bad), "worst": worst}))
sys.exit(1 if bad else 0)
