"""GPU suite: the CUDA path, called through the C ABI (fft_gpu_* / fft_auto / fftb200_*), against the
CPU oracle on identical inputs.

Bar (BASELINE.json north_star): relative L2 error <= 1e-12 against the reference's CPU output for double
precision, forward and inverse, plus forward-then-inverse round trips. The kernels read the reference's
own twiddle recurrence from host-built tables, so the match is ~3e-16 even where the reference itself is
1e-11 away from the exact DFT (N >= 2^20).
"""
import ctypes as C
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
TOL = 1e-12  # relative L2, double precision (north_star)
GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "reference_outputs.npz"))


def batch_for(n):
    return max(1, min(67, (1 << 17) // n)) + (3 if n <= 4096 else 0)


@pytest.mark.parametrize("log_n", list(range(5, 21)))
@pytest.mark.parametrize("direction", [-1, 1])
def test_pow2_parity(gpu, port, O, log_n, direction):
    n = 1 << log_n
    b = batch_for(n)
    x = port.fill(43, 0, n * b).reshape(b, n)
    y = gpu.gpu_fft_batch(x, direction)
    assert O.rel_l2(y, port.fft_batch(x, direction)) <= TOL


@pytest.mark.parametrize("log_n", [1, 2, 3, 4])
@pytest.mark.parametrize("direction", [-1, 1])
def test_tiny_sizes_are_the_true_dft(gpu, port, O, log_n, direction):
    """N in {4, 8, 16}: the reference output is wrong (missing bit reversal, fft_common.h:59-77); the GPU
    path computes the DFT, checked against the O(n^2) sum. N = 2 matches the reference."""
    n = 1 << log_n
    x = port.fill(42, 0, n * 70).reshape(70, n)
    y = gpu.gpu_fft_batch(x, direction)
    want = np.stack([port.naive_dft(r, direction) for r in x])
    assert O.rel_l2(y, want) <= TOL
    assert O.rel_l2(y, port.fft_batch(x, direction)) <= TOL  # oracle without the quirk


@pytest.mark.parametrize("log_n", [22, 24])
def test_large_single_transform(gpu, port, O, log_n):
    n = 1 << log_n
    x = port.fill(44, 0, n).reshape(1, n)
    y = gpu.gpu_fft_batch(x, -1)
    assert O.rel_l2(y, port.fft_batch(x, -1)) <= TOL
    back = gpu.gpu_fft_batch(y, 1, inplace=True)
    assert O.rel_l2(back, x) <= 1e-9  # the reference's own round trip is ~2e-10 here (twiddle recurrence)
    assert O.rel_l2(back, port.fft_batch(y, 1)) <= TOL


@pytest.mark.parametrize("n", [1 << 6, 1 << 10, 1 << 12, 1 << 13, 1 << 15, 1 << 18])
def test_in_place_equals_out_of_place(gpu, port, n):
    b = batch_for(n)
    x = port.fill(45, 0, n * b).reshape(b, n)
    assert np.array_equal(gpu.gpu_fft_batch(x, -1, inplace=True), gpu.gpu_fft_batch(x, -1, inplace=False))


@pytest.mark.parametrize("n", [3, 5, 6, 7, 12, 97, 360, 1009, 4099, 100003, 1000003])
@pytest.mark.parametrize("direction", [-1, 1])
def test_bluestein_parity(gpu, port, O, n, direction):
    b = 3 if n < 5000 else 1
    x = port.fill(46, 0, n * b).reshape(b, n)
    y = gpu.gpu_fft_batch(x, direction)
    if n <= 8:  # the reference's Bluestein is broken for m <= 16 (same missing bit reversal)
        want = np.stack([port.naive_dft(r, direction) for r in x])
    else:
        want = np.stack([port.fft(r, direction) for r in x])
    assert O.rel_l2(y, want) <= TOL


@pytest.mark.parametrize("n,batch", [(257, 1), (300, 1000), (1009, 67), (2003, 5), (2048 - 1, 300), (1500, 3)])
@pytest.mark.parametrize("direction", [-1, 1])
def test_bluestein_inside_the_pipe_kernel_is_bit_identical(gpu, port, O, n, batch, direction, monkeypatch):
    """Bluestein with padded length m = 512 .. 4096: chirp multiply + zero padding ride on the first gather of the forward
    transform, the spectral product on its stores, the final chirp multiply on the stores of the inverse (fft_pipe.cuh,
    PIPE_BLUE_FWD / INV): 2 launches instead of 5. FFTB200_NO_PIPE_BLUE=1 is the five-kernel path: same arithmetic, identical
    bits; first and last rows against the oracle; in place."""
    x = port.fill(56, 0, n * batch).reshape(batch, n)
    a = gpu.gpu_fft_batch(x, direction)
    monkeypatch.setenv("FFTB200_NO_PIPE_BLUE", "1")
    b = gpu.gpu_fft_batch(x, direction)
    monkeypatch.delenv("FFTB200_NO_PIPE_BLUE")
    assert np.array_equal(a, b)
    assert np.array_equal(gpu.gpu_fft_batch(x, direction, inplace=True), a)
    rows = sorted({0, batch - 1})
    assert O.rel_l2(a[rows], np.stack([port.fft(x[r], direction) for r in rows])) <= TOL


@pytest.mark.parametrize("n,batch", [(2049, 7), (4095, 33), (8193, 3), (10000, 37), (16384 - 1, 9), (20011, 5), (65536 + 1, 3), (100003, 2), (262144 - 5, 2), (500009, 1), (524288 - 1, 3)])
@pytest.mark.parametrize("direction", [-1, 1])
def test_bluestein_inside_the_fused_kernel_is_bit_identical(gpu, port, O, n, batch, direction, monkeypatch):
    """Bluestein with padded length m = 2^13 .. 2^20: the forward transform reads the caller's rows through a tensor map that ends
    at the last full row (zero padding = out-of-range rows, the partial row read by the kernel) and multiplies by conj(chirp) in the
    first gather; the inverse multiplies by FB in its first gather and by conj(chirp) / n on the way out, storing the first n values of
    every row from registers (fft_fused.cuh, FUSED_BLUE_FWD / INV): 2 launches instead of 5. FFTB200_FUSED_BLUE_MODE=0 is the five-kernel path around the
    same fused transforms: same arithmetic, identical bits; first and last rows against the oracle; in place. Sizes around every padded length, n a multiple of
    the row length R and not, batches that leave ragged groups."""
    x = port.fill(57, 0, n * batch).reshape(batch, n)
    a = gpu.gpu_fft_batch(x, direction)
    monkeypatch.setenv("FFTB200_FUSED_BLUE_MODE", "0")
    b = gpu.gpu_fft_batch(x, direction)
    monkeypatch.delenv("FFTB200_FUSED_BLUE_MODE")
    assert np.array_equal(a, b)
    assert np.array_equal(gpu.gpu_fft_batch(x, direction, inplace=True), a)
    rows = sorted({0, batch - 1})
    assert O.rel_l2(a[rows], np.stack([port.fft(x[r], direction) for r in rows])) <= TOL


@pytest.mark.parametrize("direction", [-1, 1])
def test_bluestein_fused_chirp_is_bit_identical(gpu, port, direction, monkeypatch):
    """Multi-pass Bluestein plans (m >= 2^21) apply the chirp factors and the spectral product inside the first / last
    tile pass (fft_tile.cuh MUL_*); FFTB200_NO_FUSED_CHIRP=1 runs them as separate elementwise kernels. Same arithmetic
    in the same order: the outputs must be identical bit for bit (batch 3: ragged against the tile grid)."""
    n, b = 1000003, 3
    x = port.fill(52, 0, n * b).reshape(b, n)
    # (without the factors riding on it the last pass of the plain transforms would run in the ring kernel, whose derived twiddles differ
    # from the tile kernel's table twiddles in the last bits: keep both runs on the tile kernel, the comparison is about the factors)
    monkeypatch.setenv("FFTB200_NO_LASTPIPE", "1")
    fused = gpu.gpu_fft_batch(x, direction)
    monkeypatch.setenv("FFTB200_NO_FUSED_CHIRP", "1")
    plain = gpu.gpu_fft_batch(x, direction)
    monkeypatch.delenv("FFTB200_NO_FUSED_CHIRP")
    assert np.array_equal(fused, plain)
    assert np.array_equal(gpu.gpu_fft_batch(x, direction, inplace=True), fused)


@pytest.mark.parametrize("n", [2, 4, 64, 1024, 1 << 14, 1 << 17, 1 << 20])
def test_r2c_parity(gpu, port, O, n):
    x = port.fill(47, 0, n).real.copy()
    got = gpu.r2c(x)
    assert got.shape == (n // 2 + 1,)
    assert O.rel_l2(got, port.r2c(x)) <= TOL


@pytest.mark.parametrize("n,batch", [(256, 9), (4096, 5), (1 << 13, 41), (1 << 15, 300), (1 << 16, 7), (1 << 18, 3), (1 << 20, 5)])
def test_r2c_batched_engine_plan(gpu, port, O, n, batch):
    """Batched r2c (the GPU batch API of SURVEY 8a8): N = 2^13 .. 2^20 run inside the fused kernel (real tile in,
    bins 0 .. n/2 out, no promote / extract passes); smaller N through the promote - c2c - extract plan."""
    import torch
    L = gpu.lib
    x = port.fill(47, 0, n * batch).real.copy().reshape(batch, n)
    plan = gpu.engine_plan(n, batch, gpu.FFTB200_R2C)
    xd = torch.from_numpy(x).cuda()
    yd = torch.zeros((batch, n // 2 + 1), dtype=torch.complex128, device="cuda")
    for _ in range(2):
        assert L.fftb200_plan_exec(plan, xd.data_ptr(), yd.data_ptr()) == 0
    rows = sorted({0, batch // 2, batch - 1})
    want = np.stack([port.r2c(x[r]) for r in rows])
    assert O.rel_l2(yd[rows].cpu().numpy(), want) <= TOL
    full = np.fft.rfft(x, axis=1)   # every row, against the accurate transform (the reference is 1e-11 away from it at 2^20)
    assert O.rel_l2(yd.cpu().numpy(), full) <= 5e-11
    L.fftb200_plan_destroy(plan)


@pytest.mark.parametrize("key", sorted(k for k in GOLD.files if k.startswith("full_")))
def test_golden_full_through_fft_auto(gpu, port, O, key):
    _, n, seed, tag = key.split("_")
    n, seed, sign = int(n), int(seed), (-1 if tag == "f" else 1)
    x = port.fill(seed, 0, n)
    got = gpu.fft_auto(x, sign)
    if n in (4, 8, 16) or (n & (n - 1) and n <= 8):
        want = port.naive_dft(x, sign)  # reference output is not the DFT at these sizes
    elif sign > 0 and n == 12:
        want = GOLD[key] / n            # reference mixed-radix inverse is unscaled (mixed_radix.c:107-124)
    else:
        want = GOLD[key]
    assert O.rel_l2(got, want) <= TOL


@pytest.mark.parametrize("key", sorted(k for k in GOLD.files if k.startswith("samp_")))
def test_golden_sampled_through_fft_auto(gpu, port, O, key):
    _, n, seed, tag = key.split("_")
    n, seed, sign = int(n), int(seed), (-1 if tag == "f" else 1)
    x = port.fill(seed, 0, n)
    y = gpu.fft_auto(x, sign)
    assert O.rel_l2(y[GOLD[f"idx_{n}"]], GOLD[key]) <= TOL
    assert abs(np.linalg.norm(y) / GOLD[f"norm_{n}_{seed}_{tag}"][0] - 1.0) <= 1e-13


def test_plan_execute_dft_and_in_place_host(gpu, port, O):
    L = gpu.lib
    n = 2048
    a, b = port.fill(1, 0, n), port.fill(2, 0, n)
    out = np.zeros(n, complex)
    plan = L.fft_plan_dft_1d(n, gpu.ptr(a), gpu.ptr(out), -1, gpu.FFT_PREFER_GPU)
    assert plan
    L.fft_execute(plan)
    assert O.rel_l2(out, port.fft(a)) <= TOL
    out2 = np.zeros(n, complex)
    L.fft_execute_dft(plan, gpu.ptr(b), gpu.ptr(out2))
    assert O.rel_l2(out2, port.fft(b)) <= TOL
    L.fft_execute(plan)  # the plan's own arrays are restored after execute_dft
    assert O.rel_l2(out, port.fft(a)) <= TOL
    L.fft_destroy_plan(plan)
    c = a.copy()
    assert L.fft_auto(gpu.ptr(c), gpu.ptr(c), n, 1) == 0  # in place, inverse (sign >= 0)
    assert O.rel_l2(c, port.fft(a, 1)) <= TOL


def test_host_batch_entry_point(gpu, port, O):
    n, b = 512, 41
    x = port.fill(9, 0, n * b).reshape(b, n)
    y = np.zeros_like(x)
    assert gpu.lib.fft_gpu_dft_1d_batch(gpu.ptr(x), gpu.ptr(y), n, b, 1) == 0
    assert O.rel_l2(y, port.fft_batch(x, 1)) <= TOL
    z = np.zeros(n, complex)
    assert gpu.lib.fft_gpu_dft_1d(gpu.ptr(x[0].copy()), gpu.ptr(z), n, -1) == 0
    assert O.rel_l2(z, port.fft(x[0])) <= TOL


def test_gpu_api_error_conventions(gpu, capfd):
    L = gpu.lib
    assert L.fft_gpu_init(gpu.FFT_GPU_CUDA) == 0 and L.fft_gpu_init(gpu.FFT_GPU_AUTO) == 0  # idempotent
    assert L.fft_gpu_init(2) == -1                                                           # Metal
    assert L.fft_gpu_get_backend() == gpu.FFT_GPU_CUDA
    assert b"B200" in L.fft_gpu_get_device_name() or L.fft_gpu_get_device_name()
    tot, av = C.c_size_t(), C.c_size_t()
    L.fft_gpu_get_memory_info(C.byref(tot), C.byref(av))
    assert tot.value >= av.value > 0
    small = L.fft_gpu_alloc(16)
    plan = L.fft_gpu_plan_1d(64, 4, -1)
    L.fft_gpu_execute(plan, small, small)  # buffer too small: reported, not executed
    assert "smaller" in capfd.readouterr().err
    L.fft_gpu_destroy_plan(plan)
    L.fft_gpu_free(small)
    p2 = L.fft_gpu_plan_2d(8, 8, -1)   # a stub in the reference (gpu/fft_gpu.c:377-394); implemented here (tests/test_gpu_apps.py)
    assert p2
    L.fft_gpu_destroy_plan(p2)
    assert L.fft_get_hardware_capabilities() & (1 << 5)  # FFT_HW_GPU_CUDA


# ---- the reference's own property tests (tests/test_all.c:64-351) re-expressed on the GPU path ----
PROP_SIZES = [2, 4, 8, 16, 32, 64, 128, 256, 512, 1024, 4096, 1 << 14, 1 << 17]


@pytest.mark.parametrize("n", PROP_SIZES)
def test_properties(gpu, port, n):
    tol = 1e-10
    imp = np.zeros(n, complex)
    imp[0] = 1
    dc = np.ones(n, complex)
    a, b = port.fill(1, 0, n), port.fill(2, 0, n)
    i = np.arange(n)
    tone = (np.sin(2 * np.pi * 3 * i / n) + 0.5 * np.cos(2 * np.pi * 7 * i / n)).astype(complex)
    X = gpu.gpu_fft_batch(np.stack([imp, dc, a, b, 2 * a + 3 * b, tone]), -1)
    assert np.max(np.abs(np.abs(X[0]) - 1)) <= tol                                   # impulse
    assert abs(X[1][0] - n) <= tol * n and np.max(np.abs(X[1][1:]), initial=0) <= tol * n  # DC
    assert np.max(np.abs(X[4] - (2 * X[2] + 3 * X[3]))) <= tol * n                   # linearity
    assert abs(np.sum(np.abs(a) ** 2) - np.sum(np.abs(X[2]) ** 2) / n) <= tol * n    # Parseval
    back = gpu.gpu_fft_batch(X[5:6], 1)
    assert np.max(np.abs(back[0] - tone)) <= 1e-9                                    # round trip


# ---- BASELINE configuration at full size, through size-independent properties ----
def test_config2_full_size_round_trip_and_samples(gpu, port, O):
    """N = 4096 x 65536 (4 GiB in, 4 GiB out), generated on the device: 64 sampled transforms against the
    oracle, then inverse of the whole result against the input (relative L2 computed on the device)."""
    import torch
    L = gpu.lib
    n, batch = 4096, 65536
    x = torch.empty((batch, n), dtype=torch.complex128, device="cuda")
    y = torch.empty_like(x)
    assert L.fftb200_fill_splitmix(x.data_ptr(), 43, 0, n * batch) == 0
    torch.cuda.synchronize()
    fwd = L.fft_gpu_plan_1d(n, batch, -1)
    inv = L.fft_gpu_plan_1d(n, batch, 1)
    assert fwd and inv
    assert L.fftb200_plan_exec(L.fftb200_engine_of(fwd), x.data_ptr(), y.data_ptr()) == 0
    rng = np.random.default_rng(0)
    rows = np.unique(np.concatenate([[0, batch - 1], rng.integers(0, batch, 62)]))
    got = y[torch.as_tensor(rows, device="cuda")].cpu().numpy()
    xin = np.stack([port.fill(43, int(r) * n, n) for r in rows])
    assert np.array_equal(xin, x[torch.as_tensor(rows, device="cuda")].cpu().numpy())  # device fill == host fill
    assert O.rel_l2(got, port.fft_batch(xin, -1)) <= TOL
    # Parseval over the whole job
    ex, ey = float((x.real ** 2 + x.imag ** 2).sum()), float((y.real ** 2 + y.imag ** 2).sum())
    assert abs(ey / (n * ex) - 1) <= 1e-12
    assert L.fftb200_plan_exec(L.fftb200_engine_of(inv), y.data_ptr(), y.data_ptr()) == 0  # in place
    err = float(torch.linalg.vector_norm(y - x) / torch.linalg.vector_norm(x))
    assert err <= TOL
    L.fft_gpu_destroy_plan(fwd)
    L.fft_gpu_destroy_plan(inv)


@pytest.mark.parametrize("log_n,batch", [(13, 5000), (14, 3001), (15, 777), (16, 1000), (17, 300), (18, 100), (19, 37), (20, 24)])
def test_fused_two_pass_many_groups(gpu, port, O, log_n, batch):
    """N = 2^14 .. 2^20 (and 2^13 under FFTB200_NO_PIPE13) run in the fused two-pass kernel (csrc/fft_fused.cuh). Batches large enough that the
    L2-resident scratch ring wraps many times and ragged against its grouping: sampled transforms against the
    oracle, Parseval over the whole job, then the in-place inverse of everything against the input."""
    import torch
    L = gpu.lib
    n = 1 << log_n
    x = torch.empty((batch, n), dtype=torch.complex128, device="cuda")
    y = torch.empty_like(x)
    assert L.fftb200_fill_splitmix(x.data_ptr(), 48, 0, n * batch) == 0
    torch.cuda.synchronize()
    fwd = L.fft_gpu_plan_1d(n, batch, -1)
    inv = L.fft_gpu_plan_1d(n, batch, 1)
    assert fwd and inv
    assert L.fftb200_plan_exec(L.fftb200_engine_of(fwd), x.data_ptr(), y.data_ptr()) == 0
    rng = np.random.default_rng(log_n)
    rows = np.unique(np.concatenate([[0, batch - 1], rng.integers(0, batch, 6)]))
    got = y[torch.as_tensor(rows, device="cuda")].cpu().numpy()
    xin = np.stack([port.fill(48, int(r) * n, n) for r in rows])
    assert O.rel_l2(got, port.fft_batch(xin, -1)) <= TOL
    ex, ey = float((x.real ** 2 + x.imag ** 2).sum()), float((y.real ** 2 + y.imag ** 2).sum())
    assert abs(ey / (n * ex) - 1) <= (1e-12 if log_n <= 16 else 1e-10)  # the reference's drifting twiddles are not unitary
    for _ in range(2):  # the same plan again: counters and scratch ring are reset per execution
        assert L.fftb200_plan_exec(L.fftb200_engine_of(fwd), x.data_ptr(), y.data_ptr()) == 0
    assert L.fftb200_plan_exec(L.fftb200_engine_of(inv), y.data_ptr(), y.data_ptr()) == 0  # in place
    err = float(torch.linalg.vector_norm(y - x) / torch.linalg.vector_norm(x))
    assert err <= (TOL if log_n <= 16 else 2e-11)  # the reference's own round trip is 1e-11 at 2^20 (twiddle recurrence)
    L.fft_gpu_destroy_plan(fwd)
    L.fft_gpu_destroy_plan(inv)


@pytest.mark.parametrize("batch", [1, 2, 3, 147, 149, 445, 5000])
def test_pipe13_one_visit_kernel(gpu, port, O, batch):
    """N = 8192 runs in fft_pipe13_kernel (csrc/fft_pipe13.cuh): the two 4096-point halves of a transform in a three-buffer
    ring, de-interleaved on the first gather, stage 13 traded between the two thread groups. Batches below, at and ragged against the grid
    (148 CTAs), so CTAs with 0, 1, 2 and many transforms and every ring phase are exercised; forward against the oracle,
    the same plan twice, then the in-place inverse of everything against the input."""
    import torch
    L = gpu.lib
    n = 1 << 13
    x = torch.empty((batch, n), dtype=torch.complex128, device="cuda")
    y = torch.empty_like(x)
    assert L.fftb200_fill_splitmix(x.data_ptr(), 50, 0, n * batch) == 0
    torch.cuda.synchronize()
    fwd = L.fft_gpu_plan_1d(n, batch, -1)
    inv = L.fft_gpu_plan_1d(n, batch, 1)
    assert fwd and inv
    assert b"P13" in L.fftb200_plan_describe(L.fftb200_engine_of(fwd))
    for _ in range(2):
        assert L.fftb200_plan_exec(L.fftb200_engine_of(fwd), x.data_ptr(), y.data_ptr()) == 0
    rng = np.random.default_rng(batch)
    rows = np.unique(np.concatenate([[0, batch - 1], rng.integers(0, batch, 8)]))
    got = y[torch.as_tensor(rows, device="cuda")].cpu().numpy()
    xin = np.stack([port.fill(50, int(r) * n, n) for r in rows])
    assert O.rel_l2(got, port.fft_batch(xin, -1)) <= TOL
    ex, ey = float((x.real ** 2 + x.imag ** 2).sum()), float((y.real ** 2 + y.imag ** 2).sum())
    assert abs(ey / (n * ex) - 1) <= 1e-12
    assert L.fftb200_plan_exec(L.fftb200_engine_of(inv), y.data_ptr(), y.data_ptr()) == 0  # in place
    err = float(torch.linalg.vector_norm(y - x) / torch.linalg.vector_norm(x))
    assert err <= TOL
    L.fft_gpu_destroy_plan(fwd)
    L.fft_gpu_destroy_plan(inv)


def test_fused_kernel_at_2_13_agrees_with_pipe13(gpu, port, O, monkeypatch):
    """FFTB200_NO_PIPE13=1 plans N = 8192 in the fused two-pass kernel (Z7+6) as before; both routes must match the oracle
    and each other to the parity bar (they differ only in the stage-8 .. 13 twiddle tables: reference recurrence vs accurate)."""
    n, batch = 1 << 13, 333
    x = port.fill(51, 0, n * batch).reshape(batch, n)
    a = gpu.gpu_fft_batch(x, -1)
    monkeypatch.setenv("FFTB200_NO_PIPE13", "1")
    b = gpu.gpu_fft_batch(x, -1)
    monkeypatch.delenv("FFTB200_NO_PIPE13")
    want = port.fft_batch(x[:4], -1)
    assert O.rel_l2(a[:4], want) <= TOL and O.rel_l2(b[:4], want) <= TOL
    assert O.rel_l2(a, b) <= 1e-13


@pytest.mark.parametrize("n", [2, 64, 1000, 1024, 4096, 4099])
def test_small_host_transforms_zero_copy_equals_staged(gpu, port, O, n, monkeypatch):
    """Host-pointer transforms of up to 128 KB with a single FFT kernel run through a device-mapped page-locked staging
    buffer (one launch, one synchronise; fft_auto on 1024 points: 36 -> 22 us); FFTB200_NO_ZEROCOPY=1 is the three-stream
    staged-copy path. Same kernels on the same values: bit-identical, c2c / Bluestein / r2c / c2r, repeated calls."""
    x = port.fill(54, 0, n)
    pow2 = n & (n - 1) == 0

    def run():
        r = [gpu.fft_auto(x, s) for s in (-1, 1)] + [gpu.fft_auto(x, -1)]
        if pow2:
            r += [gpu.r2c(x.real.copy()), gpu.c2r(gpu.r2c(x.real.copy()), n)]
        return r
    a = run()
    monkeypatch.setenv("FFTB200_NO_ZEROCOPY", "1")
    b = run()
    monkeypatch.delenv("FFTB200_NO_ZEROCOPY")
    for u, v in zip(a, b):
        assert np.array_equal(u, v)
    assert O.rel_l2(a[0], port.fft(x, -1)) <= TOL


@pytest.mark.parametrize("log_n,batch", [(22, 11), (23, 9), (24, 8)])
@pytest.mark.parametrize("direction", [-1, 1])
def test_tma_ring_last_pass_equals_tile_last_pass(gpu, port, O, log_n, batch, direction, monkeypatch):
    """N = 2^22 .. 2^24: the last pass runs in fft_lastpipe_kernel (TMA ring, the fused kernel's pass-B dataflow); FFTB200_NO_LASTPIPE=1
    keeps fft_tile_kernel<LAST>. With FFTB200_LASTPIPE_TABLE=1 the ring kernel loads every twiddle from the table like the tile kernel:
    bit-identical where the radix split is the same (2^23, 2^24), within 1e-15 otherwise. By default it loads 4 of the 15 twiddles of a
    radix-16 butterfly and derives the others as products with the table's own entries at kappa = 0 (the reference's recurrence is
    multiplicative up to rounding noise - the noise is what differs: 2.9e-14 at 2^22, 7.5e-14 at 2^23, 2.9e-13 at 2^24, the largest size
    that uses it); last transform against the oracle either way."""
    n = 1 << log_n
    x = port.fill(55, 0, n * batch).reshape(batch, n)
    y = gpu.gpu_fft_batch(x, direction)
    monkeypatch.setenv("FFTB200_NO_LASTPIPE", "1")
    z = gpu.gpu_fft_batch(x, direction)
    monkeypatch.delenv("FFTB200_NO_LASTPIPE")
    monkeypatch.setenv("FFTB200_LASTPIPE_TABLE", "1")
    w = gpu.gpu_fft_batch(x, direction)
    monkeypatch.delenv("FFTB200_LASTPIPE_TABLE")
    if log_n >= 23:
        assert np.array_equal(w, z)
    else:
        assert O.rel_l2(w, z) <= 1e-15
    d = O.rel_l2(y, z)
    print("derived vs table twiddles, 2^%d: %.2e" % (log_n, d))
    assert d <= (4e-13 if log_n == 24 else 1e-13)
    assert O.rel_l2(y[-1:], port.fft_batch(x[-1:], direction)) <= TOL


def test_two_fused_plans_run_concurrently(gpu, port, O):
    """The fused kernel's CTAs synchronise through global counters and must all be resident; it is launched
    cooperatively, so two plans enqueued back to back on their own streams must both finish with the right answer."""
    import torch
    L = gpu.lib
    n, batch = 1 << 16, 300
    x = torch.empty((2, batch, n), dtype=torch.complex128, device="cuda")
    y = torch.empty_like(x)
    assert L.fftb200_fill_splitmix(x.data_ptr(), 49, 0, 2 * n * batch) == 0
    torch.cuda.synchronize()
    plans = [L.fft_gpu_plan_1d(n, batch, -1) for _ in range(2)]
    for rep in range(3):
        for i, pl in enumerate(plans):
            assert L.fftb200_plan_exec_async(L.fftb200_engine_of(pl), x[i].data_ptr(), y[i].data_ptr()) == 0
    for pl in plans:
        assert L.fftb200_plan_sync(L.fftb200_engine_of(pl)) == 0
    for i in range(2):
        rows = [0, 17, batch - 1]
        got = y[i][rows].cpu().numpy()
        xin = np.stack([port.fill(49, (i * batch + r) * n, n) for r in rows])
        assert O.rel_l2(got, port.fft_batch(xin, -1)) <= TOL
    for pl in plans:
        L.fft_gpu_destroy_plan(pl)


def test_config3_linearity_at_2_24(gpu):
    import torch
    L = gpu.lib
    n = 1 << 24
    x = torch.empty((3, n), dtype=torch.complex128, device="cuda")
    assert L.fftb200_fill_splitmix(x.data_ptr(), 44, 0, 2 * n) == 0
    torch.cuda.synchronize()
    x[2] = 2 * x[0] + 3 * x[1]
    torch.cuda.synchronize()
    y = torch.empty_like(x)
    plan = L.fft_gpu_plan_1d(n, 3, -1)
    assert plan
    assert L.fftb200_plan_exec(L.fftb200_engine_of(plan), x.data_ptr(), y.data_ptr()) == 0
    err = float(torch.linalg.vector_norm(y[2] - (2 * y[0] + 3 * y[1])) / torch.linalg.vector_norm(y[2]))
    assert err <= TOL
    L.fft_gpu_destroy_plan(plan)


# ---- distributed transform (fft-implementation-in-c_b200/dist.py): partial plans + rank tables on one GPU, ranks on several ----
def _dist_module():
    import importlib.util
    import fftb200_loader
    spec = importlib.util.spec_from_file_location("fft_b200_dist", os.path.join(fftb200_loader.PKG_DIR, "dist.py"))
    D = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(D)
    return D


@pytest.mark.parametrize("log_n,direction", [(20, -1), (21, 1), (22, -1), (24, -1)])
def test_distributed_plan_world_1_matches_the_oracle(gpu, port, O, log_n, direction):
    """World size 1 runs the same head / tail partial plans, rank table, push kernel and peer-store epilogue (onto
    itself) as any rank of a multi-GPU job; both drivers."""
    import torch
    D = _dist_module()
    n = 1 << log_n
    x = port.fill(45, 0, n)
    want = port.fft(x, direction)
    plan = D.DistFFT(n, 1, 0, lambda *a: D.CudaBackend(gpu, *a), direction=direction)
    y = plan.execute(torch.from_numpy(x).cuda())
    torch.cuda.synchronize()
    assert O.rel_l2(y.cpu().numpy(), want) <= TOL
    plan.close()
    plan = D.DistFFTP2P(gpu, n, 1, 0, direction=direction)
    for _ in range(2):
        y = plan.execute(torch.from_numpy(x).cuda())
        torch.cuda.synchronize()
        assert O.rel_l2(y.cpu().numpy(), want) <= TOL
    plan.close()


def test_distributed_two_ranks_match_the_single_gpu_plan(gpu):
    """Two ranks over NCCL (needs two GPUs; the single-GPU box of the round-end run skips it)."""
    import subprocess
    import sys
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    import json
    for extra in ([], ["--nccl"]):   # fused P2P exchanges (the product) and the NCCL all-to-all baseline driver
        r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                            "--master-port", "29533", os.path.join(root, "tools", "dist_check.py"), "22", "24", "--inverse", "--notime"] + extra,
                           capture_output=True, text=True, timeout=600)
        assert r.returncode == 0, r.stderr[-2000:]
        errs = [json.loads(l)["rel_l2_vs_single_gpu_plan"] for l in r.stdout.replace("}{", "}\n{").splitlines() if l.startswith("{")]
        assert len(errs) == 8 and max(errs) <= TOL
