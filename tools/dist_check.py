"""Distributed transform on real GPUs (run under torchrun, one rank per GPU, or plainly for world 1):
parity of every rank's output block against the single-GPU plan of the whole transform (itself checked against
the oracle in tests/), and device-timed throughput (max over ranks).
usage: [torchrun --nproc-per-node G] python tools/dist_check.py log_n [log_n ...] [--notime] [--big]"""
import ctypes as C, json, math, os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import numpy as np
import torch
import torch.distributed as dist
import fftb200_loader

rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
local = int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
F = fftb200_loader.load(); L = F.lib
L.fftb200_set_device(local)
F.require_gpu()
import importlib.util
spec = importlib.util.spec_from_file_location("fft_b200_dist", os.path.join(fftb200_loader.PKG_DIR, "dist.py"))
D = importlib.util.module_from_spec(spec); spec.loader.exec_module(D)

logs = [int(a) for a in sys.argv[1:] if not a.startswith("-")]
for lg in logs:
    n = 1 << lg
    nloc = n // world
    for direction in ((-1, 1) if "--inverse" in sys.argv else (-1,)):
        t0 = time.time()
        if "--nccl" in sys.argv:
            plan = D.DistFFT(n, world, rank, lambda *a: D.CudaBackend(F, *a), direction=direction)
        else:
            plan = D.DistFFTP2P(F, n, world, rank, direction=direction)
        t_plan = time.time() - t0
        x = torch.empty(nloc, dtype=torch.complex128, device="cuda")
        assert L.fftb200_fill_splitmix(x.data_ptr(), 45, rank * nloc, nloc) == 0
        torch.cuda.synchronize()
        y = plan.execute(x)
        torch.cuda.synchronize()
        res = {"driver": "nccl" if "--nccl" in sys.argv else "p2p", "log_n": lg, "world": world, "rank": rank, "dir": direction, "log_m": plan.log_m, "plan_s": round(t_plan, 2), "passes": plan.describe if hasattr(plan, "describe") else plan.be.describe}
        if "--noparity" not in sys.argv and n * 16 * 3 < 120e9 and (lg < 30 or rank == 0):
          try:
            full = torch.empty(n, dtype=torch.complex128, device="cuda")
            assert L.fftb200_fill_splitmix(full.data_ptr(), 45, 0, n) == 0
            ref = torch.empty_like(full)
            p1 = L.fft_gpu_plan_1d(n, 1, direction)
            assert L.fftb200_plan_exec(L.fftb200_engine_of(p1), full.data_ptr(), ref.data_ptr()) == 0
            L.fft_gpu_destroy_plan(p1)
            mine = ref[rank * nloc:(rank + 1) * nloc]
            res["rel_l2_vs_single_gpu_plan"] = float(torch.linalg.vector_norm(y - mine) / torch.linalg.vector_norm(mine))
            del full, ref
          except Exception as e:  # the comparison is best effort at the largest size (host table + 3 full-size device arrays)
            res["parity_error"] = repr(e)[:200]
        if "--oracle" in sys.argv and direction == -1:
            # the reference oracle itself (unmodified split_radix_fft on the host, tests/golden/make_oracle_2p30.py): exact strided
            # bins of this rank's block + the random-sign sketch of the whole block (tests/sketch.py)
            sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "tests"))
            import sketch
            gold = os.path.join(os.path.dirname(__file__), "..", "tests", "golden")
            fx = os.path.join(gold, f"oracle_2p{lg}_sketch.npz")
            if os.path.exists(fx):
                z = np.load(fx)
                assert int(z["log_n"]) == lg and int(z["seed"]) == 45
                chunk = 1 << int(z["log_chunk"])
                sk, en = sketch.sketch_torch(y, first=rank * nloc)
                c0, c1 = rank * nloc // chunk, (rank + 1) * nloc // chunk
                res["oracle_sketch_rel_l2"] = sketch.rel_l2_estimate(sk, z["sketch"][c0:c1], z["energy"][c0:c1])
                res["oracle_energy_rel"] = float(abs(en.sum() - z["energy"][c0:c1].sum()) / z["energy"][c0:c1].sum())
                for name, lstep in (("strided", max(0, lg - 20)), ("strided_small", max(0, lg - 16))):
                    f2 = os.path.join(gold, f"oracle_2p{lg}_{name}.npy")
                    if os.path.exists(f2):
                        ref_bins = np.load(f2, mmap_mode="r")
                        step = 1 << lstep
                        mine = np.asarray(ref_bins[rank * nloc // step:(rank + 1) * nloc // step])
                        got = y[::step].cpu().numpy()
                        res[f"oracle_{name}_bins"] = int(mine.size)
                        res[f"oracle_{name}_rel_l2"] = float(np.linalg.norm(got - mine) / np.linalg.norm(mine))
                        res[f"oracle_{name}_max_abs"] = float(np.abs(got - mine).max())
            else:
                res["oracle"] = "fixture missing: " + os.path.basename(fx)
        if "--notime" not in sys.argv:
            ts = []
            for it in range(6):
                if world > 1: dist.barrier()
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()   # execute() orders the plan's stream after / before torch's current stream
                if "--nccl" in sys.argv: plan.execute(x, y)
                else: plan.execute(x)
                e1.record()
                torch.cuda.synchronize()
                t = torch.tensor([e0.elapsed_time(e1)], device="cuda")
                if world > 1: dist.all_reduce(t, op=dist.ReduceOp.MAX)
                if it >= 2: ts.append(float(t))
            res["ms"] = round(min(ts), 3)
            res["gflops"] = round(5 * n * lg / min(ts) * 1e-6)
            res["strict_GBps_aggregate"] = round(32 * n / min(ts) * 1e-6)
        print(json.dumps(res), flush=True)
        plan.close()
if world > 1:
    dist.destroy_process_group()
