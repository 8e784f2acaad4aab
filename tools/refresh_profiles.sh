#!/bin/bash
# Run on the GPU box (gpurun): full GPU test suite, bench line, all-configs table, ncu launch list and one --set full capture
# of the headline kernel. Outputs under gpurun_out/; tools/make_profiles.py turns them into profiles/*.md afterwards.
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py > gpurun_out/bench_line.json 2> gpurun_out/bench_err.log; tail -c 600 gpurun_out/bench_line.json
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_line.json 2>> gpurun_out/bench_err.log
python tools/configs_bench.py > gpurun_out/configs_bench.log 2>&1; tail -3 gpurun_out/configs_bench.log | cut -c1-200
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r01_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu --e2e-steps 1 > gpurun_out/ncu_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:fft_pipe -s 3 -c 1 -f -o gpurun_out/r01_pipe python bench.py --steps 1 --warmup 3 --no-cpu --no-e2e > gpurun_out/ncu_full.log 2>&1

# second-tier kernels: one --set full capture each (N = 8192 one-visit kernel, fused two-pass kernel at 2^16 and 2^20)
ncu --set full --clock-control none --import-source on -k regex:fft_pipe13 -s 2 -c 1 -f -o gpurun_out/r01_pipe13 python tools/one.py 13 > gpurun_out/ncu_pipe13.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:fft_fused -s 2 -c 1 -f -o gpurun_out/r01_fused_16 python tools/one.py 16 > gpurun_out/ncu_fused16.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:fft_fused -s 2 -c 1 -f -o gpurun_out/r01_fused_20 python tools/one.py 20 > gpurun_out/ncu_fused20.log 2>&1
python tools/blue_time.py 1000003 1 16 64 > gpurun_out/blue_time.log 2>&1
ls -la gpurun_out/ | tail -5
