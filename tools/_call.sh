set -x
mkdir -p gpurun_out
for w in config2 config1 config4 config5_yaml; do python tools/perf_frame.py $w 4 2>&1 | tail -1; done > gpurun_out/r02_perf4.log 2>&1
cat gpurun_out/r02_perf4.log
