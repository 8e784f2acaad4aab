# same-box timing of Bluestein plans, five kernels (mode 0) against the fused variants (mode 3); development tool
for n in ${NS:-3000 6000 9000 20000 40000 100000 200000 400000}; do
  m=1; while [ $m -lt $((2*n-1)) ]; do m=$((m*2)); done
  b=$((134217728 / m))
  for mode in ${MODES:-0 3}; do echo -n "n=$n m=$m b=$b mode=$mode "; FFTB200_FUSED_BLUE_MODE=$mode python tools/blue_time.py $n $b | python -c "import sys,json; d=json.loads(sys.stdin.readline()); print(d['ms_best'], d['ms_med'], d['launches'], d['plan'][:60])"; done
done
