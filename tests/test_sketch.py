"""CPU suite: the random-sign sketch used to check 2^30-point results against the reference oracle without the 16 GiB reference
(tests/sketch.py, tests/golden/make_oracle_2p30.py): numpy and torch hashes agree bit for bit, the estimator tracks the true
relative L2 error, and the committed 2^30 fixture is self-consistent with its exact strided bins."""
import os

import numpy as np
import pytest

import sketch

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def test_numpy_and_torch_hashes_agree():
    import torch
    for first in (0, 12345, (1 << 30) - 5000, (1 << 33) + 7):
        a = sketch.hash_numpy(np.arange(first, first + 5000, dtype=np.uint64))
        b = sketch.hash_torch(torch.arange(first, first + 5000, dtype=torch.int64)).numpy().view(np.uint64)
        assert np.array_equal(a, b)


@pytest.mark.parametrize("noise", [1e-15, 3e-13, 1e-9])
def test_estimator_tracks_the_true_error(noise):
    import torch
    rng = np.random.default_rng(7)
    n = 1 << 22
    x = rng.standard_normal(n) + 1j * rng.standard_normal(n)
    y = x + noise * (rng.standard_normal(n) + 1j * rng.standard_normal(n))
    skx, en = sketch.sketch_numpy(x, first=3 << 20)
    sky, en2 = sketch.sketch_torch(torch.from_numpy(y), first=3 << 20)
    true = np.linalg.norm(y - x) / np.linalg.norm(x)
    est = sketch.rel_l2_estimate(sky, skx, en)
    assert 0.5 * true <= est <= 2.0 * true or est < 1e-15
    assert np.allclose(en, en2, rtol=1e-6)   # en2 is the energy of the perturbed vector


def test_a_localised_error_is_seen():
    """One wrong chunk (e.g. a rank that wrote the wrong block) cannot hide: the sketch is per 2^20-element chunk."""
    rng = np.random.default_rng(8)
    n = 1 << 21
    x = rng.standard_normal(n) + 1j * rng.standard_normal(n)
    y = x.copy()
    y[(1 << 20) + 17] += 1e-3
    skx, en = sketch.sketch_numpy(x)
    sky, _ = sketch.sketch_numpy(y)
    assert sketch.rel_l2_estimate(sky, skx, en) == pytest.approx(1e-3 / np.linalg.norm(x), rel=1e-6)
    assert np.array_equal(sky[0], skx[0])


def test_committed_2p30_fixture_is_consistent(port):
    """The committed sketch and strided bins of the reference's 2^30-point output (seed 45) have the documented shapes; the input
    energy relation of the unnormalised transform holds (Parseval: sum |X|^2 = N sum |x|^2, x uniform in [-1, 1)^2 -> 2/3 per point)."""
    z = np.load(os.path.join(GOLD, "oracle_2p30_sketch.npz"))
    assert int(z["log_n"]) == 30 and int(z["seed"]) == 45 and int(z["log_chunk"]) == sketch.LOG_CHUNK and int(z["k"]) == sketch.K
    assert z["sketch"].shape == (1024, sketch.K) and z["energy"].shape == (1024,)
    n = float(1 << 30)
    assert z["energy"].sum() == pytest.approx(n * n * 2.0 / 3.0, rel=1e-3)
    small = np.load(os.path.join(GOLD, "oracle_2p30_strided_small.npy"))
    assert small.shape == (1 << 16,) and small.dtype == np.complex128
    big = os.path.join(GOLD, "oracle_2p30_strided.npy")   # git-ignored (16 MiB): present where the generator ran
    if os.path.exists(big):
        assert np.array_equal(np.load(big, mmap_mode="r")[::16], small)
