"""PCIe ceiling on this box (development tool): pinned H2D alone, D2H alone, both at once, per NUMA placement."""
import os, sys, time, subprocess
import torch
print(subprocess.run(["nvidia-smi", "topo", "-m"], capture_output=True, text=True).stdout)
print("cpus:", os.cpu_count(), "affinity:", len(os.sched_getaffinity(0)))
try:
    print(open("/sys/devices/system/node/online").read().strip(), "nodes online")
except OSError:
    pass
N = 1 << 30  # 1 GiB
def run(tag):
    h_in = torch.empty(N, dtype=torch.uint8).pin_memory()
    h_out = torch.empty(N, dtype=torch.uint8).pin_memory()
    d_a = torch.empty(N, dtype=torch.uint8, device="cuda"); d_b = torch.empty(N, dtype=torch.uint8, device="cuda")
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    def t(fn, reps=4):
        fn(); torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(reps): fn()
        torch.cuda.synchronize()
        return (time.perf_counter() - t0) / reps
    def up():
        with torch.cuda.stream(s1): d_a.copy_(h_in, non_blocking=True)
    def down():
        with torch.cuda.stream(s2): h_out.copy_(d_b, non_blocking=True)
    def both():
        up(); down()
    a, b, c = t(up), t(down), t(both)
    print(f"{tag}: H2D {N/a*1e-9:.1f} GB/s  D2H {N/b*1e-9:.1f} GB/s  duplex {N/c*1e-9:.1f} GB/s each way", flush=True)
run("default affinity")
ncpu = os.cpu_count()
for lo, hi in [(0, ncpu // 2), (ncpu // 2, ncpu)]:
    try:
        os.sched_setaffinity(0, range(lo, hi))
        run(f"cpus {lo}-{hi-1}")
    except OSError as e:
        print("affinity", lo, hi, e)
