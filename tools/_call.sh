mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -x -q --durations=8 2>&1 | tail -25 > gpurun_out/r02_pytest21.log
cat gpurun_out/r02_pytest21.log
L=gpurun_out/r02_config3.log; : > $L
for p in "1.5 16 64" "1.95 16 64" "3.0 8 128" "3.0 11 128" "3.0 14 64"; do
  timeout 600 python tools/bench_config3.py 10000 3840 2160 $p 2>/dev/null | tail -1 >> $L
done
cat $L
