"""Platform PCIe ceiling with ALL GPUs of the box copying at once (development tool; run under torchrun, one rank per GPU):
pinned 1 GiB H2D alone, D2H alone and both directions at once, every rank starting together after a barrier. The aggregate
duplex figure is the ceiling of bench.py's `e2e` at N GPUs (each rank moves 4 GiB up + 4 GiB down per step).
usage: torchrun --nproc-per-node G tools/pcie_bench_multi.py"""
import json, os, subprocess, time
import torch
import torch.distributed as dist

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
N = 1 << 30
h_in = torch.empty(N, dtype=torch.uint8).pin_memory()
h_out = torch.empty(N, dtype=torch.uint8).pin_memory()
d_a = torch.empty(N, dtype=torch.uint8, device="cuda"); d_b = torch.empty(N, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def up():
    with torch.cuda.stream(s1): d_a.copy_(h_in, non_blocking=True)
def down():
    with torch.cuda.stream(s2): h_out.copy_(d_b, non_blocking=True)
def both():
    up(); down()


def timed(fn, reps=4):
    fn(); torch.cuda.synchronize()
    if world > 1: dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps): fn()
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / reps
    t = torch.tensor([dt], dtype=torch.float64, device="cuda")
    if world > 1: dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return dt, float(t)


res = {}
for name, fn in (("h2d", up), ("d2h", down), ("duplex", both)):
    mine, worst = timed(fn)
    res[name] = {"rank_GBps": N / mine * 1e-9, "aggregate_GBps_each_way": world * N / worst * 1e-9}
allr = [None] * world
if world > 1: dist.all_gather_object(allr, res)
else: allr = [res]
if rank == 0:
    try:
        print(subprocess.run(["nvidia-smi", "topo", "-m"], capture_output=True, text=True).stdout)
    except Exception: pass
    print("cpus:", os.cpu_count(), "numa nodes:", open("/sys/devices/system/node/online").read().strip() if os.path.exists("/sys/devices/system/node/online") else "?")
    for name in ("h2d", "d2h", "duplex"):
        print(json.dumps({"test": name, "gpus": world, "per_rank_GBps": [round(r[name]["rank_GBps"], 1) for r in allr],
                          "aggregate_GBps_each_way": round(allr[0][name]["aggregate_GBps_each_way"], 1)}))
if world > 1: dist.destroy_process_group()
