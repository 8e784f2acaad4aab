for gt in 1 2 4; do for lag in 1 2 3; do for slots in 2 3 4; do
  [ $slots -le $lag ] && continue
  echo -n "gt=$gt lag=$lag slots=$slots "; FFTB200_FUSED_GT=$gt FFTB200_FUSED_LAG=$lag FFTB200_FUSED_SLOTS=$slots python tools/time_nb.py 24 1 20 | python -c "import sys,json; d=json.loads(sys.stdin.readline()); print(d['best'], d['med'])"
done; done; done
echo -n "default "; python tools/time_nb.py 24 1 20 | cut -c1-80
