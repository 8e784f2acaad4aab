#!/bin/bash
# compute-sanitizer over the kernels / variants added in round 2 (run under gpurun; summary -> profiles/r02_sanitizer.md)
run() { echo "== $1: $2"; timeout 280 compute-sanitizer --tool $1 $2 2>&1 | grep -E "ERROR SUMMARY|RACECHECK SUMMARY|Illegal|Invalid|hazard" | sort | uniq -c | head -5; }
for tool in memcheck racecheck; do
  run $tool "python tools/real_time.py 15 13"      # r2c on the Hermitian schedule + c2r reading half spectra, fused<8,7>
  run $tool "python tools/real_time.py 16 5"       # fused<8,8>
done
run memcheck "python tools/real_time.py 20 2"       # fused<10,10>: r2c full pass B, c2r half spectra
run memcheck "fft-implementation-in-c_b200/bin/demo_dist 22 1"   # push kernel, peer-store epilogue, peer-flag barrier (world 1)
run racecheck "fft-implementation-in-c_b200/bin/demo_dist 22 1"
run memcheck "python tools/one.py 12 500 1"         # headline kernel with the bounded wait
run memcheck "python tools/one.py n1000 40 1"       # Bluestein in the pipe kernel
