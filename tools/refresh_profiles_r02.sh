#!/bin/bash
# Run on the GPU box (gpurun): round-2 evidence. Outputs under gpurun_out/; tools/make_profiles.py turns the ncu files into profiles/*.md.
python -m pytest tests -m gpu -x -q 2>&1 | tail -3 | tee gpurun_out/r02_pytest_gpu.txt
python bench.py > gpurun_out/r02_bench_line.json 2> gpurun_out/r02_bench_err.txt; tail -c 300 gpurun_out/r02_bench_line.json; echo
python bench.py --impl reference > gpurun_out/r02_bench_reference_line.json 2>> gpurun_out/r02_bench_err.txt
python tools/cufft_bench.py > gpurun_out/r02_cufft.md 2>&1; tail -3 gpurun_out/r02_cufft.md | cut -c1-200
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu --e2e-steps 1 --no-secondary > gpurun_out/r02_ncu_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:fft_pipe -s 3 -c 1 -f -o gpurun_out/r02_pipe python bench.py --steps 1 --warmup 3 --no-cpu --no-e2e --no-secondary > gpurun_out/r02_ncu_full.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:fft_pipe13 -s 2 -c 1 -f -o gpurun_out/r02_pipe13 python tools/one.py 13 > gpurun_out/r02_ncu_pipe13.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:fft_fused -s 2 -c 1 -f -o gpurun_out/r02_fused_16 python tools/one.py 16 > gpurun_out/r02_ncu_fused16.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:fft_fused -s 2 -c 1 -f -o gpurun_out/r02_fused_20 python tools/one.py 20 > gpurun_out/r02_ncu_fused20.log 2>&1
ls -la gpurun_out/*.ncu-rep | tail -5
