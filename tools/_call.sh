mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 > gpurun_out/r02_pytest15.log
L=gpurun_out/r02_perf15.log; : > $L
for w in config2 config1 config4 config5_yaml; do python tools/perf_frame.py $w 4 2>&1 | tail -1 >> $L; done
NDT_B200_LIB=$PWD/ndt_b200/variants/libndt_b200_nopre_8.so python tools/perf_frame.py config2 4 2>&1 | tail -1 >> $L
NDT_B200_LIB=$PWD/ndt_b200/variants/libndt_b200_nopre_4.so python tools/perf_frame.py config1 4 2>&1 | tail -1 >> $L
NDT_B200_LIB=$PWD/ndt_b200/variants/libndt_b200_nopre_6.so python tools/perf_frame.py config4 4 2>&1 | tail -1 >> $L
cat gpurun_out/r02_pytest15.log $L
