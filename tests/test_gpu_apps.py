"""GPU suite, "next" rows of SURVEY.md 8f: c2r, 2-D transforms, FFT convolution / correlation - the CUDA path
through the C ABI (fft_plan_c2r_1d, fft_plan_dft_2d, fft_gpu_plan_2d, fft_gpu_dft_2d, fft_gpu_convolution, ...)
against the oracle's restatements of the reference's own callers (pinned in tests/test_oracle_apps.py) and against
the committed outputs of the unmodified reference functions (tests/golden/reference_apps.npz).
Tolerance: relative L2 <= 1e-12 (north_star, double precision)."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
TOL = 1e-12
GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "reference_apps.npz"))
SEED2 = 1 << 20


@pytest.mark.parametrize("n", [2, 4, 16, 64, 1024, 4096, 1 << 13, 1 << 16, 1 << 20])
def test_c2r_parity_and_round_trip(gpu, port, O, n):
    x = port.fill(73, 0, n).real.copy()
    half = gpu.r2c(x)
    back = gpu.c2r(half, n)
    assert back.shape == (n,)
    assert O.rel_l2(back, port.c2r(half, n)) <= TOL       # same operator as the oracle's inverse
    assert O.rel_l2(back, x) <= (1e-12 if n <= 1 << 16 else 1e-10)   # r2c -> c2r round trip (reference recurrence error above 2^16)


@pytest.mark.parametrize("n,batch", [(256, 7), (8192, 5), (1 << 15, 3)])
def test_c2r_batched_engine_plan(gpu, port, O, n, batch):
    import torch
    L = gpu.lib
    x = port.fill(74, 0, n * batch).real.copy().reshape(batch, n)
    half = np.stack([port.r2c(r) for r in x])
    plan = gpu.engine_plan(n, batch, gpu.FFTB200_C2R, direction=1)
    hd = torch.from_numpy(half).cuda()
    yd = torch.zeros((batch, n), dtype=torch.float64, device="cuda")
    assert L.fftb200_plan_exec(plan, hd.data_ptr(), yd.data_ptr()) == 0
    want = np.stack([port.c2r(h, n) for h in half])
    assert O.rel_l2(yd.cpu().numpy(), want) <= TOL
    L.fftb200_plan_destroy(plan)


@pytest.mark.parametrize("n,batch", [(512, 1), (512, 2), (512, 1187), (1024, 5), (1024, 16), (2048, 300), (4096, 1), (4096, 2), (4096, 301)])
def test_real_variants_of_the_pipe_kernel(gpu, port, O, n, batch, monkeypatch):
    """Batched r2c / c2r of 512 .. 4096 points run in fft_pipe_kernel's real variants (reals / half spectra read and written by
    the kernel itself); FFTB200_NO_PIPE_REAL=1 is promote - c2c - extract (r2c) and extend - inverse c2c - real parts (c2r).
    Same operators (two real transforms share one complex transform in the kernel: exact to rounding), equal to the oracle. Batches ragged
    against the tiles of 8192 reals, odd batches leave a pair half empty."""
    import torch
    L = gpu.lib
    x = port.fill(75, 0, n * batch).real.copy().reshape(batch, n)
    xd = torch.from_numpy(x).cuda()

    def run():
        fwd = gpu.engine_plan(n, batch, gpu.FFTB200_R2C)
        inv = gpu.engine_plan(n, batch, gpu.FFTB200_C2R, direction=1)
        desc = L.fftb200_plan_describe(fwd) + L.fftb200_plan_describe(inv)
        hd = torch.zeros((batch, n // 2 + 1), dtype=torch.complex128, device="cuda")
        yd = torch.zeros((batch, n), dtype=torch.float64, device="cuda")
        for _ in range(2):
            assert L.fftb200_plan_exec(fwd, xd.data_ptr(), hd.data_ptr()) == 0
            assert L.fftb200_plan_exec(inv, hd.data_ptr(), yd.data_ptr()) == 0
        L.fftb200_plan_destroy(fwd)
        L.fftb200_plan_destroy(inv)
        return hd.cpu().numpy(), yd.cpu().numpy(), desc
    h1, y1, d1 = run()
    assert b"no promote" in d1 and b"no extension" in d1
    monkeypatch.setenv("FFTB200_NO_PIPE_REAL", "1")
    h2, y2, d2 = run()
    monkeypatch.delenv("FFTB200_NO_PIPE_REAL")
    assert b"hermitian extension" in d2
    # the kernel transforms two real rows as ONE complex transform (z = x_a + i x_b, separated at the end; two half spectra packed as
    # X_a + i X_b on the way back): exact to rounding, so the two paths agree to the last bits - and the c2r input is the r2c output of
    # each path, hence the looser second bound
    assert O.rel_l2(h1, h2) <= 2e-15 and O.rel_l2(y1, y2) <= 4e-15
    rows = sorted({0, batch // 2, batch - 1})
    assert O.rel_l2(h1[rows], np.stack([port.r2c(x[r]) for r in rows])) <= TOL
    assert O.rel_l2(y1, x) <= TOL


@pytest.mark.parametrize("n,batch", [(1 << 14, 1), (1 << 14, 77), (1 << 15, 40), (1 << 16, 9), (1 << 17, 5), (1 << 18, 3), (1 << 19, 2), (1 << 20, 2)])
def test_c2r_fused_kernel_stores_real_parts(gpu, port, O, n, batch, monkeypatch):
    """c2r of 2^14 .. 2^20 points: the fused inverse kernel stages and stores only the real parts (fft_fused.cuh, C2R);
    FFTB200_NO_FUSED_C2R=1 keeps the separate real-part pass. Same arithmetic: bit-identical; first / last rows vs the oracle."""
    import torch
    L = gpu.lib
    x = port.fill(76, 0, n * batch).real.copy().reshape(batch, n)
    half = np.fft.rfft(x, axis=1)
    hd = torch.from_numpy(half).cuda()

    def run():
        plan = gpu.engine_plan(n, batch, gpu.FFTB200_C2R, direction=1)
        desc = L.fftb200_plan_describe(plan)
        yd = torch.zeros((batch, n), dtype=torch.float64, device="cuda")
        for _ in range(2):
            assert L.fftb200_plan_exec(plan, hd.data_ptr(), yd.data_ptr()) == 0
        L.fftb200_plan_destroy(plan)
        return yd.cpu().numpy(), desc
    y1, d1 = run()
    assert b"storing the real parts" in d1
    monkeypatch.setenv("FFTB200_NO_FUSED_C2R", "1")
    y2, d2 = run()
    monkeypatch.delenv("FFTB200_NO_FUSED_C2R")
    assert b"+ real parts" in d2
    assert O.rel_l2(y1, y2) <= 2e-15   # (the fused path reads the half spectrum and transforms half of the pass-A columns: same values to the last bits)
    rows = sorted({0, batch - 1})
    assert O.rel_l2(y1[rows], np.stack([port.c2r(half[r], n) for r in rows])) <= TOL


SHAPES = [(64, 128), (256, 64), (8, 32), (1024, 128), (2, 2), (1, 64), (64, 1), (32, 4096), (4096, 32), (512, 512),
          (2048, 1024), (16, 16), (128, 4)]


@pytest.mark.parametrize("rows,cols", SHAPES)
@pytest.mark.parametrize("sign", [-1, 1])
def test_fft2d_parity(gpu, port, O, rows, cols, sign):
    x = port.fill(58, 0, rows * cols).reshape(rows, cols)
    want = port.fft2d(x, sign)   # rows then columns with the 1-D oracle, inverse scaled once (header contract)
    for api in ("plan", "gpu", "dft"):
        got = gpu.fft2d(x, sign, api=api)
        assert O.rel_l2(got, want) <= TOL, api


@pytest.mark.parametrize("key", [k for k in sorted(GOLD.files) if k.startswith("img_") and k.endswith("_f")])
def test_fft2d_against_the_reference_outputs(gpu, port, O, key):
    _, rows, cols, seed, _ = key.split("_")
    rows, cols, seed = int(rows), int(cols), int(seed)
    if rows in (4, 8, 16) or cols in (4, 8, 16):
        pytest.skip("the reference's 1-D transform is wrong at n in {4, 8, 16} (fft_common.h:59-77); covered vs the oracle")
    x = port.fill(seed, 0, rows * cols).reshape(rows, cols)
    fwd, inv = gpu.fft2d(x, -1), gpu.fft2d(x, 1)
    base = key[:-2]
    if base + "_idx" in GOLD.files:
        idx = GOLD[base + "_idx"]
        assert O.rel_l2(fwd.ravel()[idx], GOLD[base + "_f"]) <= TOL
        assert O.rel_l2(inv.ravel()[idx] / (rows * cols), GOLD[base + "_i"]) <= TOL   # the reference scales twice (image_fft.c:63-71)
    else:
        assert O.rel_l2(fwd, GOLD[base + "_f"]) <= TOL
        assert O.rel_l2(inv / (rows * cols), GOLD[base + "_i"]) <= TOL


@pytest.mark.parametrize("rows,cols", [(3, 5), (12, 100), (97, 64), (64, 97), (100, 100)])
def test_fft2d_any_shape(gpu, port, O, rows, cols):
    """Shapes that are not powers of two (the reference's fft_2d exits there): Bluestein rows / columns."""
    x = port.fill(59, 0, rows * cols).reshape(rows, cols)
    got = gpu.fft2d(x, -1)
    assert O.rel_l2(got, np.fft.fft2(x)) <= 1e-11
    assert O.rel_l2(gpu.fft2d(got, 1), x) <= 1e-11


def test_fft2d_both_column_paths_agree(gpu, port, O, monkeypatch):
    x = port.fill(60, 0, 256 * 512).reshape(256, 512)
    a = gpu.fft2d(x, -1, api="gpu")                 # strided column kernels
    monkeypatch.setenv("FFTB200_2D_TURN", "1")
    b = gpu.fft2d(x, -1, api="gpu")                 # transposes around contiguous column transforms
    assert O.rel_l2(a, b) <= 1e-13
    assert O.rel_l2(a, port.fft2d(x, -1)) <= TOL


@pytest.mark.parametrize("key", [k for k in sorted(GOLD.files) if k.startswith("conv_")])
def test_convolution_parity(gpu, port, O, key):
    _, nx, nh, seed = key.split("_")
    nx, nh, seed = int(nx), int(nh), int(seed)
    x, h = port.fill(seed, 0, nx), port.fill(seed, SEED2, nh)
    got = gpu.convolution(x, h)
    assert got.shape == (nx + nh - 1,)
    assert O.rel_l2(got, GOLD[key]) <= TOL              # the unmodified reference's fft_convolution
    assert O.rel_l2(got, port.convolution(x, h)) <= TOL


@pytest.mark.parametrize("n", [256, 4096, 1000, 1 << 16])
def test_circular_convolution_parity(gpu, port, O, n):
    x, h = port.fill(53, 0, n), port.fill(53, SEED2, n)
    got = gpu.circular_convolution(x, h)
    assert O.rel_l2(got, port.circular_convolution(x, h)) <= TOL
    key = f"circ_{n}_53"
    if key in GOLD.files:
        assert O.rel_l2(got, GOLD[key]) <= TOL


@pytest.mark.parametrize("n", [100, 1024, 3000, 1 << 17])
def test_correlation_parity(gpu, port, O, n):
    seed = {100: 55, 1024: 56, 3000: 57}.get(n, 62)
    x, y = port.fill(seed, 0, n), port.fill(seed, SEED2, n)
    xc, ac = gpu.cross_correlation(x, y), gpu.autocorrelation(x)
    assert O.rel_l2(xc, port.cross_correlation(x, y)) <= TOL
    assert O.rel_l2(ac, port.autocorrelation(x)) <= TOL
    if f"xcorr_{n}_{seed}" in GOLD.files:
        assert O.rel_l2(xc, GOLD[f"xcorr_{n}_{seed}"]) <= TOL   # the unmodified reference's cross_correlation_fft
        assert O.rel_l2(ac, GOLD[f"acorr_{n}_{seed}"]) <= TOL


def test_transpose_helper(gpu, port):
    import torch
    L = gpu.lib
    for rows, cols, batch in [(1, 1, 1), (33, 65, 3), (1000, 7, 2), (64, 4096, 1)]:
        x = torch.randn(batch, rows, cols, dtype=torch.complex128, device="cuda")
        y = torch.empty(batch, cols, rows, dtype=torch.complex128, device="cuda")
        torch.cuda.synchronize()
        assert L.fftb200_transpose(y.data_ptr(), x.data_ptr(), rows, cols, batch, None) == 0
        torch.cuda.synchronize()
        assert torch.equal(y, x.transpose(1, 2).contiguous())


def test_new_entry_points_reject_bad_arguments(gpu):
    L = gpu.lib
    buf = np.zeros(16, dtype=np.complex128)
    assert not L.fft_plan_c2r_1d(12, gpu.ptr(buf), gpu.ptr(buf), 0)       # power of two only, like r2c
    assert not L.fft_plan_c2r_1d(0, gpu.ptr(buf), gpu.ptr(buf), 0)
    assert not L.fft_plan_dft_2d(0, 4, gpu.ptr(buf), gpu.ptr(buf), -1, 0)
    assert not L.fft_plan_dft_2d(4, 4, None, gpu.ptr(buf), -1, 0)
    assert not L.fft_gpu_plan_2d(-1, 4, -1)
    assert L.fft_gpu_dft_2d(None, gpu.ptr(buf), 4, 4, -1) == -1
    assert L.fft_gpu_convolution(None, 4, gpu.ptr(buf), 4, gpu.ptr(buf)) == -1
    assert L.fft_gpu_cross_correlation(gpu.ptr(buf), gpu.ptr(buf), 0, gpu.ptr(buf)) == -1


# ---- the callers: C programs written against the public headers, and the reference's own demo ----
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "fft-implementation-in-c_b200", "bin")


def _run(path, timeout=600):
    import subprocess
    assert os.path.exists(path), path + " is not built (python __graft_entry__.py)"
    return subprocess.run([path], capture_output=True, text=True, timeout=timeout)


def test_demo_drop_in_program(gpu):
    r = _run(os.path.join(BIN, "demo_drop_in"))
    assert r.returncode == 0, r.stdout + r.stderr
    assert "all checks passed" in r.stdout and "FAILED" not in r.stdout


def test_distributed_transform_from_plain_c(gpu):
    """programs/demo_dist.c: fftb200_dist_create / exec_async / sync / destroy from C99, one thread per GPU, the all-gather callback
    a pthread barrier; every block against the single-GPU plan. Runs with every power-of-two GPU count the box offers."""
    import subprocess
    exe = os.path.join(BIN, "demo_dist")
    assert os.path.exists(exe)
    g = 1
    while g <= gpu.lib.fftb200_device_count() and g <= 8:
        r = subprocess.run([exe, "22" if g == 1 else "24", str(g)], capture_output=True, text=True, timeout=600)
        assert r.returncode == 0 and "PASS" in r.stdout, r.stdout + r.stderr
        g *= 2


def test_benchmark_program_follows_the_reference_protocol(gpu):
    """benchmarks/benchmark_all.c protocol (sizes, iterations, rand() input, PASS iff reconstruction <= 1e-10)."""
    r = _run(os.path.join(BIN, "benchmark_all_gpu"))
    assert r.returncode == 0, r.stdout + r.stderr
    assert "ALL ROWS PASS" in r.stdout
    assert r.stdout.count("PASS (recon") == 8 * 3


def test_reference_demo_runs_unmodified_on_this_library(gpu):
    """examples/demo_v2_features.c of the reference, compiled with the reference's own headers (oracle/Makefile),
    linked against libfft_b200.so: plans, fft_execute, fft_auto, the GPU demo and cleanup all go through the B200 path."""
    exe = os.path.join(ROOT, "oracle", "_ref", "ref_demo_v2_features")
    if not os.path.exists(exe):
        pytest.skip("oracle/_ref/ref_demo_v2_features not built (needs /root/reference at build time)")
    r = _run(exe)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "GPU Device: NVIDIA" in r.stdout and "NVIDIA CUDA" in r.stdout
    assert "No GPU available" not in r.stdout


def test_wisdom_lists_planned_shapes_and_reimports(gpu):
    L = gpu.lib
    import ctypes as C
    gpu.fft_auto(np.ones(2048, complex))
    p = L.fft_export_wisdom_to_string()
    text = C.string_at(p).decode()
    assert text.startswith("# FFT Wisdom v2.0.0\n")          # the reference's header line (fft_auto.c:420)
    assert "plan 2048 1 -1 0" in text
    assert L.fft_import_wisdom_from_string(text.encode()) == 1
    assert L.fft_import_wisdom_from_string(b"garbage") == 0
    assert L.fft_import_wisdom_from_string(None) == 0
