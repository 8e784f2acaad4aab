"""Distributed 1-D c2c transform: N = 2^L points block-distributed over G = 2^g GPUs, one process per GPU.

SURVEY.md 8(e) / BASELINE config 4 (N = 2^30 over 8 B200). The reference has no multi-device code; this is the
four-step split of ITS algorithm (radix-2 DIT, algorithms/core/radix2_dit.c:59-120) N = R * M, so the result is
the reference's result (late-stage twiddles come from the reference recurrence, host/ref_twiddle.c):

  input   rank s holds x[s N/G, (s+1) N/G) = rows t in its range of the row-major [M][R] view x[r + R t]
  T0      all-to-all: rank g gets the columns r in [g R/G, (g+1) R/G) of every row     -> A_g[t][r_loc]
  head    the first log2 M DIT stages = M-point transforms along t (partial plan, stride R/G) -> [k][r_loc]
  T1      all-to-all: rank g gets the rows k in [g M/G, (g+1) M/G) for every r         -> B_g[k_loc][r]
  tail    the last log2 R stages along r with the twiddles T[s][k + M q] of the rank's k range   -> [q][k_loc]
  T2      all-to-all: rank h gets q in [h R/G, (h+1) R/G) for every k                 -> X_h[q_loc][k] = X[k + M q]
  output  natural order, block-distributed like the input

Message shape per ordered GPU pair and exchange: N / G^2 complex doubles (256 MiB at N = 2^30, G = 8). Two drivers:

  DistFFTP2P (the product)  a thin binding of the C99 host API (include/fftb200_dist.h, host/fft_dist.c): the exchanges are
      FUSED into the kernels over NVSwitch peer memory: T0 is one push kernel, T1 / T2 are the final scatter of the last
      head / tail pass, which stores straight into the peers' exchange buffers (CUDA IPC mappings) in the layout the next
      local pass wants (csrc/fft_tile.cuh: peer_ptr). No NCCL data movement, no separate transposes; the phases are
      separated by the library's own peer-flag barrier kernel (fftb200_barrier_*), torch.distributed only all-gathers the
      IPC handles at plan time.
  DistFFT (baseline, and the CPU-testable statement of the algebra)  each exchange is one NCCL all_to_all_single
      plus one local block permute (fftb200_permute_bac). It is written against a small backend interface so that
      the same index algebra runs on CPU tensors with the gloo backend (tests/test_dist_gloo.py; numpy stands in
      for the kernels there).
The local passes are the engine's Stockham tile kernels through the C ABI (fftb200_plan_create_partial).
"""
import ctypes as C
import math

import numpy as np


def _feasible(cnt):
    """can `cnt` stages be split into tile-kernel passes of 6..9 stages?"""
    return any(6 * k <= cnt <= 9 * k for k in range(1, 5))


def choose_split(log_total, log_world):
    """log2 M for the head pass: as balanced as the pass sizes allow; M / G >= 16 and R / G >= 16."""
    best = None
    for lm in range(log_world + 4, log_total - log_world - 3):
        lr = log_total - lm
        if not (_feasible(lm) and _feasible(lr)):
            continue
        if best is None or abs(lm - lr) < abs(best - (log_total - best)) or (abs(lm - lr) == abs(best - (log_total - best)) and lm > best):
            best = lm
    if best is None:
        raise ValueError(f"no head/tail split for 2^{log_total} points over 2^{log_world} ranks")
    return best


class PlanDesc(C.Structure):  # struct fftb200_plan_desc (include/fftb200.h)
    _fields_ = [("n", C.c_int), ("batch", C.c_int), ("direction", C.c_int), ("kind", C.c_int),
                ("twiddles", C.c_void_p), ("table_n", C.c_int), ("chirp", C.c_void_p),
                ("twiddles_accurate", C.c_void_p), ("accurate_n", C.c_int), ("flags", C.c_uint)]


class CudaBackend:
    """Local passes and permutes on the GPU through the C ABI; buffers are torch complex128 CUDA tensors."""

    def __init__(self, F, n_total, world, rank, log_m, direction):
        import torch
        self.torch, self.F, L = torch, F, F.lib
        lt, lw = int(math.log2(n_total)), int(math.log2(world))
        self.nloc = n_total // world
        # head: stages [0, log_m) of the local array, standard reference table
        tab = L.fftb200_host_twiddles(1 << log_m)
        if not tab:
            raise RuntimeError("fftb200_host_twiddles failed")
        d = PlanDesc(self.nloc, 1, direction, F.FFTB200_C2C, tab, 1 << log_m, None, None, 0, 0)
        self.head = C.c_void_p()
        if L.fftb200_plan_create_partial(C.byref(self.head), C.byref(d), 0, log_m, 0, 1.0) != 0:
            raise RuntimeError("head plan: " + L.fftb200_last_error().decode())
        # tail: stages [log_m - log_world, log_total - log_world) with the rank's share of the late-stage tables
        self._ttab = np.empty(self.nloc - 1, dtype=np.complex128)
        if L.fftb200_host_twiddles_dist(self._ttab.ctypes.data, lt, lw, rank, log_m) != 0:
            raise RuntimeError("fftb200_host_twiddles_dist failed")
        d2 = PlanDesc(self.nloc, 1, direction, F.FFTB200_C2C, self._ttab.ctypes.data, self.nloc, None, None, 0, 0)
        self.tail = C.c_void_p()
        if L.fftb200_plan_create_partial(C.byref(self.tail), C.byref(d2), log_m - lw, lt - log_m, 1, 1.0 / n_total) != 0:
            raise RuntimeError("tail plan: " + L.fftb200_last_error().decode())
        self._ttab = None  # uploaded
        self.s_head = torch.cuda.ExternalStream(L.fftb200_plan_stream(self.head))
        self.s_tail = torch.cuda.ExternalStream(L.fftb200_plan_stream(self.tail))
        self.describe = (L.fftb200_plan_describe(self.head).decode(), L.fftb200_plan_describe(self.tail).decode())

    def empty(self):
        return self.torch.empty(self.nloc, dtype=self.torch.complex128, device="cuda")

    def stream(self):
        return self.torch.cuda.stream(self.s_head)

    def order_streams(self, before, tensors):
        """x and the result live on torch's current stream, the passes run on the head plan's stream: order the two."""
        cur = self.torch.cuda.current_stream()
        if before:
            self.s_head.wait_stream(cur)
        else:
            cur.wait_stream(self.s_head)   # also orders any later free / reuse of x and out on the caller's stream

    def permute_bac(self, dst, src, A, B, Cc):
        if self.F.lib.fftb200_permute_bac(dst.data_ptr(), src.data_ptr(), A, B, Cc, self.s_head.cuda_stream) != 0:
            raise RuntimeError(self.F.lib.fftb200_last_error().decode())

    def run_head(self, dst, src):
        if self.F.lib.fftb200_plan_exec_async(self.head, src.data_ptr(), dst.data_ptr()) != 0:
            raise RuntimeError(self.F.lib.fftb200_last_error().decode())

    def run_tail(self, dst, src):
        t = self.torch
        ev = t.cuda.Event()
        ev.record(self.s_head)
        self.s_tail.wait_event(ev)
        if self.F.lib.fftb200_plan_exec_async(self.tail, src.data_ptr(), dst.data_ptr()) != 0:
            raise RuntimeError(self.F.lib.fftb200_last_error().decode())
        ev2 = t.cuda.Event()
        ev2.record(self.s_tail)
        self.s_head.wait_event(ev2)

    def all_to_all(self, dst, src, group):
        import torch.distributed as dist
        dist.all_to_all_single(self.torch.view_as_real(dst), self.torch.view_as_real(src), group=group)

    def close(self):
        L = self.F.lib
        for p in (self.head, self.tail):
            if p:
                L.fftb200_plan_destroy(p)
        self.head = self.tail = None


class DistFFT:
    """plan = DistFFT(n_total, world, rank, backend); y_local = plan.execute(x_local)"""

    def __init__(self, n_total, world, rank, backend_factory, direction=-1, log_m=None, group=None):
        lt, lw = int(math.log2(n_total)), int(math.log2(world))
        if (1 << lt) != n_total or (1 << lw) != world:
            raise ValueError("n_total and world must be powers of two")
        self.n, self.world, self.rank, self.group = n_total, world, rank, group
        self.log_m = choose_split(lt, lw) if log_m is None else log_m
        self.M, self.R = 1 << self.log_m, 1 << (lt - self.log_m)
        self.be = backend_factory(n_total, world, rank, self.log_m, direction)
        self.b0, self.b1 = self.be.empty(), self.be.empty()

    def execute(self, x, out=None):
        be, G = self.be, self.world
        Ml, Rl = self.M // G, self.R // G
        sync = getattr(be, "order_streams", None)
        if out is None:
            out = be.empty()
        if sync:
            sync(before=True, tensors=(x, out))   # the plan's stream waits for the producer of x (and of out)
        with be.stream():
            if G > 1:
                be.permute_bac(self.b0, x, Ml, G, Rl)             # [t_loc][g][r_loc] -> [g][t_loc][r_loc]
                be.all_to_all(self.b1, self.b0, self.group)       # -> [s][t_loc][r_loc] = A[t][r_loc]
                src = self.b1
            else:
                src = x
            be.run_head(self.b0, src)                             # -> [k][r_loc] = [h][k_loc][r_loc]
            if G > 1:
                be.all_to_all(self.b1, self.b0, self.group)       # -> [s][k_loc][r_loc]
                be.permute_bac(self.b0, self.b1, G, Ml, Rl)       # -> [k_loc][s][r_loc] = B[k_loc][r]
            be.run_tail(self.b1, self.b0)                         # -> [q][k_loc] = [h][q_loc][k_loc]
            if G > 1:
                be.all_to_all(self.b0, self.b1, self.group)       # -> [s][q_loc][k_loc]
                be.permute_bac(out, self.b0, G, Rl, Ml)           # -> [q_loc][s][k_loc] = X[k + M q], natural order
            else:
                out.copy_(self.b1)
        if sync:
            sync(before=False, tensors=(x, out))  # the caller's stream waits for the result
        return out

    def close(self):
        self.be.close()


class _DevView:
    """A device pointer owned by the library, exposed to torch through __cuda_array_interface__ (complex128, n elements)."""

    def __init__(self, ptr, n):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": "<c16", "data": (int(ptr), False), "version": 2}

    def tensor(self):
        import torch
        return torch.as_tensor(self, device="cuda")


_ALLGATHER = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t)


class DistFFTP2P:
    """The product path: a binding of the C99 host API fftb200_dist_* (include/fftb200_dist.h, host/fft_dist.c). The
    exchanges are fused into the kernels (P2P stores over NVLink / NVSwitch), the phases are separated by the library's own
    peer-flag barrier; torch.distributed only serves the plan-time all-gather of the IPC handles (the callback below).

    plan = DistFFTP2P(F, n_total, world, rank, direction); y = plan.execute(x_local)
    y is a view of a buffer owned by the plan (natural order, this rank's block). It is ordered after the transform on
    torch's current stream and is OVERWRITTEN by the next execute() of any rank - consume or copy it before calling again.
    """

    def __init__(self, F, n_total, world, rank, direction=-1, log_m=None, group=None):
        import torch
        import torch.distributed as dist
        self.torch, self.F, self.group = torch, F, group
        L = F.lib
        lt = int(math.log2(n_total))
        if (1 << lt) != n_total or (world & (world - 1)):
            raise ValueError("n_total and world must be powers of two")
        self.n, self.world, self.rank, self.nloc = n_total, world, rank, n_total // world
        L.fftb200_dist_create.argtypes = [C.POINTER(C.c_void_p), C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _ALLGATHER, C.c_void_p]
        L.fftb200_dist_exec_async.argtypes = [C.c_void_p, C.c_void_p, C.POINTER(C.c_void_p)]
        L.fftb200_dist_sync.argtypes = [C.c_void_p]
        L.fftb200_dist_stream.argtypes = [C.c_void_p]
        L.fftb200_dist_stream.restype = C.c_void_p
        L.fftb200_dist_log_m.argtypes = [C.c_void_p]
        L.fftb200_dist_describe.argtypes = [C.c_void_p]
        L.fftb200_dist_describe.restype = C.c_char_p
        L.fftb200_dist_destroy.argtypes = [C.c_void_p]
        L.fftb200_dist_destroy.restype = None

        def allgather(ctx, send, recv, nbytes):
            try:
                dev = "cuda" if dist.get_backend(group) == "nccl" else "cpu"
                mine = torch.frombuffer(bytearray(C.string_at(send, nbytes)), dtype=torch.uint8).to(dev)
                parts = [torch.empty(nbytes, dtype=torch.uint8, device=dev) for _ in range(world)]
                dist.all_gather(parts, mine, group=group)
                C.memmove(recv, torch.cat(parts).cpu().numpy().tobytes(), nbytes * world)
                return 0
            except Exception:  # noqa: BLE001 - reported through the return code of the C call
                return -1
        self._cb = _ALLGATHER(allgather)   # kept alive: the plan calls it again when it is destroyed
        self.h = C.c_void_p()
        if L.fftb200_dist_create(C.byref(self.h), lt, world, rank, direction, int(log_m or 0), self._cb, None) != 0:
            raise RuntimeError("fftb200_dist_create: " + L.fftb200_last_error().decode())
        self.log_m = L.fftb200_dist_log_m(self.h)
        self.describe = L.fftb200_dist_describe(self.h).decode()
        self.s = torch.cuda.ExternalStream(L.fftb200_dist_stream(self.h))

    def stream(self):
        return self.torch.cuda.stream(self.s)

    def execute(self, x):
        t, L = self.torch, self.F.lib
        cur = t.cuda.current_stream()
        self.s.wait_stream(cur)            # x was produced on the caller's stream
        out = C.c_void_p()
        if L.fftb200_dist_exec_async(self.h, x.data_ptr(), C.byref(out)) != 0:
            raise RuntimeError("fftb200_dist_exec_async: " + L.fftb200_last_error().decode())
        cur.wait_stream(self.s)            # the result (and the last read of x) is ordered before whatever the caller enqueues next,
                                           # so x needs no record_stream (which would outlive the plan's stream and fail at free time)
        return _DevView(out.value, self.nloc).tensor()

    def close(self):
        if self.h:
            self.torch.cuda.synchronize()
            self.F.lib.fftb200_dist_destroy(self.h)   # collective (meets the other ranks through the callback)
            self.h = None
