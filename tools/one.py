"""Run one plan a few times (for ncu). usage: one.py log_n|nN [batch] [reps] [direction]   (n1000003: that length, Bluestein)"""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import fftb200_loader
F = fftb200_loader.load(); L = F.lib
F.require_gpu()
a1 = sys.argv[1]
n = int(a1[1:]) if a1.startswith("n") else 1 << int(a1)
lg = max(1, n.bit_length() - 1)
batch = int(sys.argv[2]) if len(sys.argv) > 2 else (1 << 28) >> lg
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
d = int(sys.argv[4]) if len(sys.argv) > 4 else -1
tot = n * batch
m_in = L.fft_gpu_alloc(tot); m_out = L.fft_gpu_alloc(tot)
L.fftb200_fill_splitmix(L.fftb200_devptr_of(m_in), 43, 0, tot)
plan = L.fft_gpu_plan_1d(n, batch, d)
for _ in range(reps): L.fft_gpu_execute(plan, m_in, m_out)
print(L.fftb200_plan_describe(L.fftb200_engine_of(plan)).decode())
