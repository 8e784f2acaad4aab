#!/bin/bash
# compute-sanitizer over the kernels changed late in round 2: on-demand tile hand-out (pipe, pipe13, last-pass ring), Bluestein inside the
# fused kernel, TMEM variant of the 8192-point kernel (run under gpurun; summary -> profiles/r02_sanitizer.md)
run() { echo "== $1: $2"; timeout 280 compute-sanitizer --tool $1 $2 2>&1 | grep -E "ERROR SUMMARY|RACECHECK SUMMARY|Illegal|Invalid|hazard" | sort | uniq -c | head -5; }
for tool in memcheck racecheck; do
  run $tool "python tools/one.py 12 1500 2"        # headline kernel, tiles handed out by the counter (1500 tiles over 148 CTAs)
  run $tool "python tools/one.py 10 2003 2"        # ragged last tile
  run $tool "python tools/one.py 13 700 2"         # 8192 points, transforms handed out
  run $tool "python tools/one.py 22 8 1"           # column-mode head + last-pass ring kernel with the counter
  run $tool "python tools/one.py n10000 9 1"       # Bluestein inside the fused kernel (m = 2^15)
  run $tool "python tools/one.py n3000 21 1"       # m = 8192
done
run memcheck "python tools/one.py n300000 2 1"      # m = 2^20
FFTB200_PIPE13_TMEM=1 run memcheck "python tools/one.py 13 300 2"   # tensor-memory variant
