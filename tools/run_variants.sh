#!/bin/bash
# run_variants.sh [workload] [reps]: time every library under ndt_b200/variants plus the default build
wl=${1:-config2}; reps=${2:-5}
python tools/perf_frame.py $wl $reps 2>&1 | tail -1
for f in ndt_b200/variants/libndt_b200_*.so; do
  NDT_B200_LIB=$PWD/$f timeout 120 python tools/perf_frame.py $wl $reps 2>&1 | tail -1
done
