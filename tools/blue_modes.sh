for n in 9000 20000 40000 100000 200000 400000 1000000; do
  m=16384; while [ $m -lt $((2*n-1)) ]; do m=$((m*2)); done
  b=$((134217728 / m)); [ $m -gt 1048576 ] && continue
  for mode in 0 1 2 3; do echo -n "n=$n m=$m b=$b mode=$mode "; FFTB200_FUSED_BLUE_MODE=$mode python tools/blue_time.py $n $b | python -c "import sys,json; d=json.loads(sys.stdin.readline()); print(d['ms_best'], d['ms_med'])"; done
done
