"""Distributed transform (fft-implementation-in-c_b200/dist.py), host-side logic on CPU: two gloo ranks run
the same exchange / permute algebra as the GPU path, with numpy standing in for the local CUDA passes, and the
gathered result must be the DFT. Also pins the rank-specific late-stage twiddle tables (ref_twiddle.c) to the
reference recurrence."""
import math
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


class NumpyBackend:
    """Same interface as dist.CudaBackend; exact twiddles, CPU tensors."""

    def __init__(self, n_total, world, rank, log_m, direction):
        import torch
        self.torch = torch
        self.n, self.G, self.rank, self.M = n_total, world, rank, 1 << log_m
        self.R = n_total // self.M
        self.nloc = n_total // world
        self.sign = -1.0 if direction < 0 else 1.0
        self.inv = direction > 0

    def empty(self):
        return self.torch.empty(self.nloc, dtype=self.torch.complex128)

    def stream(self):
        import contextlib
        return contextlib.nullcontext()

    def permute_bac(self, dst, src, A, B, Cc):
        dst.copy_(src.view(A, B, Cc).permute(1, 0, 2).reshape(-1))

    def run_head(self, dst, src):
        a = src.numpy().reshape(self.M, self.R // self.G)
        f = np.fft.fft(a, axis=0) if not self.inv else np.fft.ifft(a, axis=0) * self.M
        dst.copy_(self.torch.from_numpy(np.ascontiguousarray(f).reshape(-1)))

    def run_tail(self, dst, src):
        Ml = self.M // self.G
        b = src.numpy().reshape(Ml, self.R)
        k = (self.rank * Ml + np.arange(Ml))[:, None]
        r = np.arange(self.R)[None, :]
        b = b * np.exp(self.sign * 2j * np.pi * (k * r % self.n) / self.n)
        f = np.fft.fft(b, axis=1) if not self.inv else np.fft.ifft(b, axis=1) * self.R / self.n
        dst.copy_(self.torch.from_numpy(np.ascontiguousarray(f.T).reshape(-1)))   # [q][k_loc]

    def all_to_all(self, dst, src, group):
        import torch.distributed as dist
        dist.all_to_all_single(self.torch.view_as_real(dst), self.torch.view_as_real(src), group=group)

    def close(self):
        pass


def _worker(rank, world, port, log_n, direction, q):
    import torch
    import torch.distributed as dist
    import importlib.util
    spec = importlib.util.spec_from_file_location("fft_b200_dist", os.path.join(ROOT, "fft-implementation-in-c_b200", "dist.py"))
    D = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(D)
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    n = 1 << log_n
    rng = np.random.default_rng(7)
    x = rng.standard_normal(n) + 1j * rng.standard_normal(n)
    plan = D.DistFFT(n, world, rank, NumpyBackend, direction=direction)
    nloc = n // world
    y = plan.execute(torch.from_numpy(x[rank * nloc:(rank + 1) * nloc].copy()))
    want = np.fft.fft(x) if direction < 0 else np.fft.ifft(x)
    err = np.linalg.norm(y.numpy() - want[rank * nloc:(rank + 1) * nloc]) / np.linalg.norm(want)
    q.put((rank, float(err), plan.log_m))
    dist.destroy_process_group()


@pytest.mark.parametrize("world,log_n,direction", [(2, 14, -1), (2, 15, 1), (4, 16, -1)])
def test_distributed_exchange_algebra_on_gloo(world, log_n, direction):
    import torch.multiprocessing as mp
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, log_n, direction, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, err, log_m in res:
        assert err <= 1e-13, (rank, err, log_m)


@pytest.mark.parametrize("log_total,log_world,log_m", [(14, 1, 7), (16, 2, 8), (18, 3, 9), (13, 0, 6)])
def test_rank_tables_hold_the_reference_late_stage_twiddles(log_total, log_world, log_m):
    import fftb200_loader
    F = fftb200_loader.load()
    n, G, M = 1 << log_total, 1 << log_world, 1 << log_m
    full = F.host_twiddles(n)   # reference recurrence, entry (s, j) at 2^(s-1) - 1 + j (bit-exact vs the oracle: test_oracle.py)
    Ml = M // G
    for rank in range(G):
        t = np.zeros(n // G - 1, dtype=np.complex128)
        assert F.lib.fftb200_host_twiddles_dist(t.ctypes.data, log_total, log_world, rank, log_m) == 0
        for s in range(log_m + 1, log_total + 1):
            sl = s - log_world
            loc = t[(1 << (sl - 1)) - 1:(1 << sl) - 1]
            glob = full[(1 << (s - 1)) - 1:(1 << s) - 1]
            jj = np.arange(loc.size)
            kloc, qq = jj % Ml, jj // Ml
            assert np.array_equal(loc, glob[kloc + rank * Ml + M * qq]), (rank, s)
    assert F.lib.fftb200_host_twiddles_dist(None, log_total, log_world, 0, log_m) == -1


def test_choose_split_is_feasible():
    import importlib.util
    spec = importlib.util.spec_from_file_location("fft_b200_dist", os.path.join(ROOT, "fft-implementation-in-c_b200", "dist.py"))
    D = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(D)
    for lt in range(20, 31):
        for lw in range(0, 4):
            lm = D.choose_split(lt, lw)
            assert D._feasible(lm) and D._feasible(lt - lm) and lm - lw >= 4 and lt - lm - lw >= 4
