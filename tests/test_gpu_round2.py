"""GPU suite, round 2 additions: every planner branch of csrc/fft_plan.cu (build_passes) against the oracle up to 2^28
points, the host plan cache (LRU, concurrent shapes), the in-library multi-device fan-out, one-point transforms, the
out-of-place rule of the single-kernel real transforms, cleanup with live plans.

Bar as everywhere: relative L2 <= 1e-12 against the reference's CPU arithmetic (oracle/fft_oracle.c, pinned to the compiled
reference by tests/test_oracle.py)."""
import ctypes as C
import os
import threading
import time

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
TOL = 1e-12


def _describe(gpu, n, batch, direction=-1):
    plan = gpu.engine_plan(n, batch, gpu.FFTB200_C2C, direction)
    d = gpu.lib.fftb200_plan_describe(plan).decode()
    gpu.lib.fftb200_plan_destroy(plan)
    return d


# (log_n, batch, substring the plan description must contain): one case per branch of build_passes
BRANCHES = [
    (21, 1, "F7"),            # 2^21 below 8 transforms: three tile passes
    (21, 8, "Zc8+8"),         # 2^21 from 8 transforms: fused column head + L5
    (22, 1, "F"),             # a single 2^22: three tile passes
    (22, 2, "Zc8+8"),         # fused column head + L6
    (23, 1, "Zc8+8"),
    (25, 1, "Zc8+8"),         # + L9
    (26, 1, "F9"),            # 2^26, 2^27: three tile passes of 8..9 stages
    (27, 1, "F9"),
    (28, 1, "F7"),            # four tile passes
]


@pytest.mark.parametrize("log_n,batch,tag", BRANCHES)
def test_planner_branches_large_sizes_match_the_oracle(gpu, port, O, log_n, batch, tag):
    n = 1 << log_n
    desc = _describe(gpu, n, batch)
    assert tag in desc, desc
    x = port.fill(44, 0, n * batch).reshape(batch, n)
    y = gpu.gpu_fft_batch(x, -1)
    rows = sorted({0, batch - 1})
    for r in rows:   # first and last transform against the oracle (seconds each on the host up to 2^28)
        want = port.fft_batch(x[r:r + 1], -1)
        assert O.rel_l2(y[r:r + 1], want) <= TOL, (log_n, batch, r, desc)
    if log_n <= 25:
        back = gpu.gpu_fft_batch(y, 1, inplace=True)
        assert O.rel_l2(back, x) <= 1e-9     # the reference's own round trip drifts with its twiddle recurrence
        assert O.rel_l2(back[:1], port.fft_batch(y[:1], 1)) <= TOL


def test_one_point_transforms_are_the_identity(gpu):
    x = np.array([[1.5 - 2.0j], [0.25 + 4.0j], [-3.0 + 0.5j]])
    for d in (-1, 1):
        assert np.array_equal(gpu.gpu_fft_batch(x, d), x)
        assert np.array_equal(gpu.gpu_fft_batch(x, d, inplace=True), x)
    assert np.array_equal(gpu.fft_auto(x[0], -1), x[0])
    assert np.array_equal(gpu.fft2d(x[:1], -1), x[:1])


def _stats(gpu):
    b, h = C.c_longlong(), C.c_longlong()
    gpu.lib.fftb200_host_cache_stats(C.byref(b), C.byref(h))
    return b.value, h.value


def test_plan_cache_keeps_both_directions_and_several_shapes(gpu, port, O):
    """The convolution pattern: forward and inverse transforms of the same length alternate (and a Bluestein length in
    between). After the first round nothing is rebuilt."""
    x1 = port.fill(60, 0, 1024)
    x2 = port.fill(61, 0, 1009)
    for _ in range(2):   # warm: four shapes enter the cache
        for x in (x1, x2):
            for s in (-1, 1):
                gpu.fft_auto(x, s)
    b0, h0 = _stats(gpu)
    reps = 50
    t0 = time.perf_counter()
    for _ in range(reps):
        y = gpu.fft_auto(x1, -1)
        z = gpu.fft_auto(y, 1)
    dt = (time.perf_counter() - t0) / (2 * reps)
    for _ in range(5):
        gpu.fft_auto(x2, -1); gpu.fft_auto(x2, 1)
    b1, h1 = _stats(gpu)
    assert b1 == b0, "a cached shape was rebuilt"
    assert h1 - h0 == 2 * reps + 10
    assert O.rel_l2(z, x1) <= 1e-12
    assert O.rel_l2(gpu.fft_auto(x2, -1), port.fft(x2, -1)) <= TOL
    print(f"fft_auto(1024) alternating directions: {dt * 1e6:.1f} us per call (ctypes overhead included)")
    assert dt < 200e-6


def test_two_threads_with_different_shapes_run_concurrently(gpu, port, O):
    shapes = [(4096, 512), (1 << 15, 64)]
    xs = [port.fill(62 + i, 0, n * b).reshape(b, n) for i, (n, b) in enumerate(shapes)]
    outs = [np.empty_like(x) for x in xs]
    errs = []

    def work(i):
        n, b = shapes[i]
        for _ in range(6):
            rc = gpu.lib.fft_gpu_dft_1d_batch(gpu.ptr(xs[i]), gpu.ptr(outs[i]), n, b, -1)
            if rc != 0:
                errs.append((i, rc))

    th = [threading.Thread(target=work, args=(i,)) for i in range(2)]
    [t.start() for t in th]
    [t.join() for t in th]
    assert not errs
    for x, y in zip(xs, outs):
        rows = [0, x.shape[0] - 1]
        assert O.rel_l2(y[rows], port.fft_batch(x[rows], -1)) <= TOL


def test_same_shape_from_two_threads(gpu, port, O):
    """A busy cache entry is not shared: the second thread gets a plan of its own."""
    n, b = 2048, 256
    x = port.fill(64, 0, n * b).reshape(b, n)
    want = port.fft_batch(x[:2], -1)
    outs = [np.empty_like(x) for _ in range(3)]
    rcs = [None] * 3

    def work(i):
        for _ in range(4):
            rcs[i] = gpu.lib.fft_gpu_dft_1d_batch(gpu.ptr(x), gpu.ptr(outs[i]), n, b, -1)

    th = [threading.Thread(target=work, args=(i,)) for i in range(3)]
    [t.start() for t in th]
    [t.join() for t in th]
    assert rcs == [0, 0, 0]
    for y in outs:
        assert O.rel_l2(y[:2], want) <= TOL
        assert np.array_equal(y, outs[0])


@pytest.mark.parametrize("n,batch", [(4096, 4099), (1 << 16, 301), (1009, 4000)])
def test_multi_device_fan_out_behind_the_batch_entry_point(gpu, port, O, n, batch):
    """fftb200_host_set_gpus(G): fft_gpu_dft_1d_batch cuts the batch into G ranges, one host thread + plan + staging ring per
    device (devices wrap around when the box has fewer, so a 1-GPU box runs the same code with both ranges on device 0)."""
    x = port.fill(65, 0, n * batch).reshape(batch, n)
    one = np.empty_like(x)
    gpu.lib.fftb200_host_set_gpus(1)
    assert gpu.lib.fft_gpu_dft_1d_batch(gpu.ptr(x), gpu.ptr(one), n, batch, -1) == 0
    try:
        for g in (2, 3):
            gpu.lib.fftb200_host_set_gpus(g)
            assert gpu.lib.fftb200_host_get_gpus() == g
            out = np.zeros_like(x)
            assert gpu.lib.fft_gpu_dft_1d_batch(gpu.ptr(x), gpu.ptr(out), n, batch, -1) == 0
            assert np.array_equal(out, one)
    finally:
        gpu.lib.fftb200_host_set_gpus(1)
    rows = [0, batch // 2, batch - 1]
    want = np.stack([port.fft(x[r], -1) for r in rows])
    assert O.rel_l2(one[rows], want) <= TOL


@pytest.mark.parametrize("n", [512, 4096, 1 << 14, 1 << 17])
def test_single_kernel_real_transforms_reject_overlapping_buffers(gpu, port, n):
    """r2c / c2r plans whose kernel reads packed rows (pipe variants 512 .. 4096, fused r2c) run out of place: an aliased call fails
    with an error instead of racing."""
    import torch
    L = gpu.lib
    batch = 40
    buf = torch.zeros(batch * (n + 2), dtype=torch.complex128, device="cuda")   # room for either side of either transform
    other = torch.zeros(batch * (n + 2), dtype=torch.complex128, device="cuda")
    p = gpu.engine_plan(n, batch, gpu.FFTB200_R2C)
    assert L.fftb200_plan_exec(p, buf.data_ptr(), buf.data_ptr()) != 0
    assert b"out of place" in L.fftb200_last_error()
    assert L.fftb200_plan_exec(p, buf.data_ptr(), buf.data_ptr() + 8 * batch * n - 16) != 0   # the last input double overlaps the first bin
    assert L.fftb200_plan_exec(p, buf.data_ptr(), other.data_ptr()) == 0
    L.fftb200_plan_destroy(p)
    p = gpu.engine_plan(n, batch, gpu.FFTB200_C2R, 1)   # pipe variants up to 4096, the fused kernel reading half spectra above
    assert L.fftb200_plan_exec(p, buf.data_ptr(), buf.data_ptr()) != 0
    L.fftb200_plan_destroy(p)


def test_c2r_in_place_where_it_is_allowed(gpu, port, O):
    """c2r outside the single-kernel sizes (here 2^21: three tile passes) goes through the plan's work array: in place is fine."""
    import torch
    L = gpu.lib
    n, batch = 1 << 21, 3
    x = port.fill(66, 0, n * batch).real.copy().reshape(batch, n)
    half = np.fft.rfft(x, axis=1)
    p = gpu.engine_plan(n, batch, gpu.FFTB200_C2R, 1)
    buf = torch.zeros(batch * n, dtype=torch.complex128, device="cuda")   # room for either side
    buf[:half.size] = torch.from_numpy(half.ravel()).cuda()
    assert L.fftb200_plan_exec(p, buf.data_ptr(), buf.data_ptr()) == 0
    got = torch.view_as_real(buf).ravel()[:batch * n].cpu().numpy().reshape(batch, n)
    assert O.rel_l2(got, x) <= 1e-10   # against numpy's accurate inverse: the reference's own arithmetic is 2e-11 from it here
    L.fftb200_plan_destroy(p)


@pytest.mark.parametrize("n,batch", [(1 << 13, 1), (1 << 13, 130), (1 << 14, 1), (1 << 14, 77), (1 << 15, 40), (1 << 16, 9), (1 << 17, 5), (1 << 18, 3), (1 << 19, 2), (1 << 20, 3)])
def test_c2r_reads_the_half_spectrum_inside_the_fused_kernel(gpu, port, O, n, batch, monkeypatch):
    """c2r of 2^14 .. 2^20 points: the fused inverse kernel loads the half spectrum itself - each pass-A tile is two boxes, its own columns and
    the mirrored ones, and the first gather reads the mirrored half backwards (fft_fused.cuh, C2R + HERM). FFTB200_C2R_HERMITIAN=0 keeps the
    separate c2r_expand pass over a full-length work array. Nothing is approximated (the extension X[N - j] = conj X[j] IS the definition of
    c2r); pass A transforms only the columns c <= R/2 and pass B rebuilds the rest with one accurate-table multiply, so the two paths agree to
    the last bits. Rows against the oracle."""
    import torch
    L = gpu.lib
    x = port.fill(83, 0, n * batch).real.copy().reshape(batch, n)
    half = np.fft.rfft(x, axis=1)
    half[:, 0] += 0.25j * np.arange(1, batch + 1)     # imaginary parts in bin 0 / Nyquist: both paths must treat them alike
    half[:, -1] -= 0.5j
    hd = torch.from_numpy(half).cuda()

    def run():
        plan = gpu.engine_plan(n, batch, gpu.FFTB200_C2R, direction=1)
        desc = L.fftb200_plan_describe(plan).decode()
        yd = torch.full((batch, n), float("nan"), dtype=torch.float64, device="cuda")
        for _ in range(2):
            assert L.fftb200_plan_exec(plan, hd.data_ptr(), yd.data_ptr()) == 0, L.fftb200_last_error()
        L.fftb200_plan_destroy(plan)
        return yd.cpu().numpy(), desc
    y1, d1 = run()
    assert "half spectrum in" in d1, d1
    monkeypatch.setenv("FFTB200_C2R_HERMITIAN", "0")
    y2, d2 = run()
    monkeypatch.delenv("FFTB200_C2R_HERMITIAN")
    assert "hermitian extension" in d2, d2
    assert np.isfinite(y1).all()
    # pass A runs on the columns c <= R/2 only; pass B rebuilds the others as w_M^-k conj(Y[R - c][k]) - one more rounding on half of its inputs
    assert O.rel_l2(y1, y2) <= 2e-15
    rows = sorted({0, batch - 1})
    assert O.rel_l2(y1[rows], np.stack([port.c2r(half[r], n) for r in rows])) <= TOL


def test_cleanup_keeps_the_tables_of_live_plans(gpu, port, O):
    """fft_gpu_cleanup drops the device's cached twiddle tables; a plan created before it still owns a reference."""
    L = gpu.lib
    n, b = 1 << 16, 5
    x = port.fill(67, 0, n * b).reshape(b, n)
    plan = L.fft_gpu_plan_1d(n, b, -1)
    m = L.fft_gpu_alloc(n * b)
    L.fft_gpu_copy_h2d(m, gpu.ptr(x), n * b)
    L.fft_gpu_cleanup()
    gpu.require_gpu()
    other = gpu.gpu_fft_batch(port.fill(68, 0, 1 << 18).reshape(1, -1), -1)   # new tables get uploaded meanwhile
    assert other.shape == (1, 1 << 18)
    L.fft_gpu_execute(plan, m, m)
    out = np.empty_like(x)
    L.fft_gpu_copy_d2h(gpu.ptr(out), m, n * b)
    assert O.rel_l2(out, port.fft_batch(x, -1)) <= TOL
    L.fft_gpu_destroy_plan(plan)
    L.fft_gpu_free(m)


def test_plan_follows_its_device(gpu, port, O):
    """A plan created on device A still runs on A after the caller moved on with fft_gpu_set_device (needs 2 GPUs)."""
    L = gpu.lib
    if L.fftb200_device_count() < 2:
        pytest.skip("needs 2 GPUs")
    n, b = 4096, 64
    x = port.fill(69, 0, n * b).reshape(b, n)
    plan = L.fft_gpu_plan_1d(n, b, -1)
    m = L.fft_gpu_alloc(n * b)
    L.fft_gpu_copy_h2d(m, gpu.ptr(x), n * b)
    assert L.fft_gpu_set_device(1) == 0
    L.fft_gpu_execute(plan, m, m)
    assert L.fft_gpu_set_device(0) == 0
    out = np.empty_like(x)
    L.fft_gpu_copy_d2h(gpu.ptr(out), m, n * b)
    assert O.rel_l2(out, port.fft_batch(x, -1)) <= TOL
    L.fft_gpu_destroy_plan(plan)
    L.fft_gpu_free(m)


@pytest.mark.parametrize("n,batch", [(1 << 13, 33), (1 << 14, 1), (1 << 14, 77), (1 << 15, 40), (1 << 16, 9), (1 << 17, 5), (1 << 18, 3), (1 << 19, 2),
                                     (1 << 20, 3)])
def test_hermitian_r2c_in_the_fused_kernel(gpu, port, O, n, batch, monkeypatch):
    """r2c of 2^13 .. 2^16 points: pass B of the fused kernel transforms only the columns k <= M/2 and writes the bins of the columns
    M - k as conjugates (fft_fused.cuh, HERM). FFTB200_NO_FUSED_R2C=1 is the promote -> full c2c -> extract plan: the
    directly computed bins (j mod M <= M/2) are the same arithmetic, to the last bits (pass A packs two real columns into one complex transform); the mirrored ones are conj X[j] where the reference has
    its own X[N - j] - equal only to the accuracy of its twiddle recurrence, which is why sizes above 2^16 keep the full pass B (measured
    mismatch 7.9e-13 at 2^17, 1.8e-12 at 2^18, 7e-12 at 2^20: profiles/r02_real.md). Every size against the oracle at the 1e-12 bar."""
    import torch
    L = gpu.lib
    x = port.fill(81, 0, n * batch).real.copy().reshape(batch, n)
    xd = torch.from_numpy(x).cuda()

    def run():
        plan = gpu.engine_plan(n, batch, gpu.FFTB200_R2C)
        desc = L.fftb200_plan_describe(plan).decode()
        yd = torch.full((batch, n // 2 + 1), float("nan"), dtype=torch.complex128, device="cuda")
        for _ in range(2):
            assert L.fftb200_plan_exec(plan, xd.data_ptr(), yd.data_ptr()) == 0, L.fftb200_last_error()
        L.fftb200_plan_destroy(plan)
        return yd.cpu().numpy(), desc
    y1, d1 = run()
    herm = (1 << 13) <= n <= (1 << 16)
    assert ("columns k <= M/2" in d1) == herm, d1
    assert np.isfinite(y1.view(np.float64)).all(), "a bin was not written"
    monkeypatch.setenv("FFTB200_NO_FUSED_R2C", "1")
    y2, d2 = run()
    monkeypatch.delenv("FFTB200_NO_FUSED_R2C")
    assert "no promote" not in d2
    if herm:
        lm = (int(np.log2(n)) + 1) // 2
        k = np.arange(n // 2 + 1) % (1 << lm)
        direct = k <= (1 << lm) // 2
        # directly computed bins: the promoted path's arithmetic up to the packed pass A (two real columns per complex transform, exact to
        # rounding); mirrored bins: conj X[j] for the reference's X[N - j]
        assert O.rel_l2(y1[:, direct], y2[:, direct]) <= 2e-15
        assert O.rel_l2(y1, y2) <= 5e-13
        # forcing the schedule off gives the full pass B on every column: the promoted path's result to the last bits
        monkeypatch.setenv("FFTB200_R2C_HERMITIAN", "0")
        y3, d3 = run()
        monkeypatch.delenv("FFTB200_R2C_HERMITIAN")
        assert "columns k <= M/2" not in d3 and O.rel_l2(y3, y2) <= 2e-15
    elif n > (1 << 16):
        # full pass B on every column; the pass-A rows above M/2 are read as conjugates of the rows M - k (Hermitian to rounding: stages m <= 1024
        # with conjugate-symmetric tables), so the two paths agree to the last bits, not bit for bit
        assert O.rel_l2(y1, y2) <= 2e-15
    rows = sorted({0, batch - 1})
    assert O.rel_l2(y1[rows], np.stack([port.r2c(x[r]) for r in rows])) <= TOL


@pytest.mark.parametrize("n", [1, 3, 12, 360, 1000, 1009, 5000, 100003])
def test_r2c_of_any_length(gpu, port, O, n):
    """fft_plan_r2c_1d plans every length like the reference (fft_auto.c:391-403: promote, then a c2c plan, which routes lengths that
    are not powers of two to Bluestein): promotion, Bluestein c2c and the cut to bins 0 .. n/2 all run on the device."""
    x = port.fill(82, 0, max(n, 2)).real.copy()[:n]
    got = gpu.r2c(x)
    assert got.shape == (n // 2 + 1,)
    if n <= 8 and n > 2:   # the reference's Bluestein is broken for m <= 16 (missing bit reversal): the true DFT is the yardstick there
        want = port.naive_dft(x.astype(np.complex128), -1)[: n // 2 + 1]
    else:
        want = port.r2c(x)
    assert O.rel_l2(got, want) <= TOL


@pytest.mark.parametrize("log_n", [26, 28, 30])
def test_single_gpu_plan_against_the_reference_oracle_sketch(gpu, log_n):
    """One transform of 2^26 / 2^28 / 2^30 points on ONE GPU (three / four tile passes) against the UNMODIFIED reference's split_radix_fft
    on the same input (seed 45): the host oracle ran once (minutes, 16 GiB at 2^30; tests/golden/make_oracle_2p30.py) and left exact strided
    bins plus a random-sign sketch of the whole output (tests/sketch.py), so the full vector is checked here without the 16 GiB reference."""
    import os
    import sys
    import torch
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    import sketch
    gold = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    z = np.load(os.path.join(gold, "oracle_2p%d_sketch.npz" % log_n))
    L = gpu.lib
    n = 1 << log_n
    free_b, total_b = torch.cuda.mem_get_info()
    if free_b < 3.2 * 16 * n + (2 << 30):
        pytest.skip("needs %d GiB of device memory" % int(3.2 * 16 * n / 2 ** 30 + 2))
    x = torch.empty(n, dtype=torch.complex128, device="cuda")
    assert L.fftb200_fill_splitmix(x.data_ptr(), int(z["seed"]), 0, n) == 0
    plan = L.fft_gpu_plan_1d(n, 1, -1)
    assert plan, L.fftb200_last_error()
    assert L.fftb200_plan_exec(L.fftb200_engine_of(plan), x.data_ptr(), x.data_ptr()) == 0, L.fftb200_last_error()
    L.fft_gpu_destroy_plan(plan)
    sk, en = sketch.sketch_torch(x, first=0)
    est = sketch.rel_l2_estimate(sk, z["sketch"], z["energy"])
    assert est <= TOL, est
    assert abs(en.sum() - z["energy"].sum()) <= 1e-12 * z["energy"].sum()
    step = 1 << (log_n - 16)
    want = np.load(os.path.join(gold, "oracle_2p%d_strided_small.npy" % log_n))
    got = x[::step].cpu().numpy()
    assert np.linalg.norm(got - want) / np.linalg.norm(want) <= TOL
    del x
    torch.cuda.empty_cache()
    L.fftb200_host_tables_release()   # the 2^30 host table is 16 GiB ...
    L.fft_gpu_cleanup()               # ... and so is the device's cached copy (freed with its last reference)
    gpu.require_gpu()


def test_random_shapes_against_numpy(gpu):
    """A bounded run of tools/fuzz.py (every kind of plan at random n / batch / direction / in-place against numpy's accurate transform):
    catches gross errors - a wrong tile, a race, a ragged batch - that fixed-size parity tests can miss. 25 s, fixed seed."""
    import json
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(root, "tools", "fuzz.py"), "25", "7"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    last = [l for l in r.stdout.splitlines() if l.startswith("{")][-1]
    res = json.loads(last)
    assert res.get("bad", []) == [] and res.get("runs", 0) >= 20, last


@pytest.mark.parametrize("kind,n,batch", [("r2c", 512, 4099), ("c2r", 512, 4099), ("r2c", 4096, 1031), ("c2r", 4096, 1031), ("r2c", 1 << 13, 517), ("c2r", 1 << 13, 517),
                                          ("r2c", 1 << 15, 259), ("c2r", 1 << 15, 259), ("r2c", 1 << 18, 33), ("c2r", 1 << 18, 33)])
def test_real_transforms_are_bit_stable_over_many_executions(gpu, port, kind, n, batch):
    """The real variants hand buffers around through mbarriers, counters and delayed refills (pair exchange in the pipe kernel, mirrored loads and
    half-column rebuilds in the fused kernel): 40 executions of a ragged batch must give the same bits every time (a race shows up as a sporadic
    difference long before it shows up as a wrong answer)."""
    import torch
    L = gpu.lib
    x = port.fill(84, 0, n * batch).real.copy().reshape(batch, n)
    if kind == "r2c":
        src = torch.from_numpy(x).cuda()
        dst = torch.zeros((batch, n // 2 + 1), dtype=torch.complex128, device="cuda")
        plan = gpu.engine_plan(n, batch, gpu.FFTB200_R2C)
    else:
        src = torch.from_numpy(np.fft.rfft(x, axis=1)).cuda()
        dst = torch.zeros((batch, n), dtype=torch.float64, device="cuda")
        plan = gpu.engine_plan(n, batch, gpu.FFTB200_C2R, direction=1)
    assert L.fftb200_plan_exec(plan, src.data_ptr(), dst.data_ptr()) == 0
    first = dst.clone()
    for it in range(40):
        dst.zero_()
        assert L.fftb200_plan_exec(plan, src.data_ptr(), dst.data_ptr()) == 0, L.fftb200_last_error()
        assert torch.equal(torch.view_as_real(dst) if kind == "r2c" else dst, torch.view_as_real(first) if kind == "r2c" else first), it
    L.fftb200_plan_destroy(plan)


@pytest.mark.parametrize("n", [512, 1024, 4096])
def test_tile_hand_out_covers_every_tile_once(gpu, n):
    """fft_pipe_kernel takes its first three tiles per CTA by position and the rest from a global counter that the last CTA resets
    (csrc/fft_pipe.cuh "Tile order"). Batches around the multiples of the grid (1 .. 3 x 148 tiles and beyond, ragged last tiles), both
    directions, several executions of the same plan in a row: every row against numpy, and the same bits every time - a tile handed out
    twice, never, or after a counter that was not reset shows up as a wrong or a stale row."""
    import torch
    L = gpu.lib
    nt = 4096 // n
    sms = torch.cuda.get_device_properties(0).multi_processor_count
    rng = np.random.default_rng(5)
    batches = [1, nt + 1, 2 * nt, sms * nt - 1, sms * nt, sms * nt + 1, 3 * sms * nt - 1, 3 * sms * nt, 3 * sms * nt + 1, 4 * sms * nt + 3, 11 * sms * nt + 5]
    for batch in batches:
        x = (rng.standard_normal((batch, n)) + 1j * rng.standard_normal((batch, n)))
        src = torch.from_numpy(x).cuda()
        dst = torch.empty_like(src)
        for direction in (-1, 1):
            plan = gpu.engine_plan(n, batch, gpu.FFTB200_C2C, direction)
            want = np.fft.fft(x, axis=1) if direction < 0 else np.fft.ifft(x, axis=1)
            first = None
            for it in range(4):
                dst.fill_(float("nan"))
                assert L.fftb200_plan_exec(plan, src.data_ptr(), dst.data_ptr()) == 0, L.fftb200_last_error()
                got = dst.cpu().numpy()
                if first is None:
                    first = got
                    err = np.linalg.norm(got - want, axis=1) / np.linalg.norm(want, axis=1)
                    assert np.all(err <= TOL), (n, batch, direction, int(np.argmax(err)), float(err.max()))
                else:
                    assert np.array_equal(got, first), (n, batch, direction, it)
            L.fftb200_plan_destroy(plan)
