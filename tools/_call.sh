mkdir -p gpurun_out
L=gpurun_out/r02_perf13.log; : > $L
for w in config2 config1 config4 config5_yaml; do python tools/perf_frame.py $w 4 2>&1 | tail -1 >> $L; done
M=gpu__time_duration.sum,smsp__sass_thread_inst_executed_op_dadd_pred_on.sum,smsp__sass_thread_inst_executed_op_dmul_pred_on.sum,smsp__sass_thread_inst_executed_op_dfma_pred_on.sum,smsp__thread_inst_executed_pipe_fp64_pred_on.sum,smsp__inst_executed_pipe_fp64.sum,smsp__inst_executed.sum,sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_elapsed,launch__registers_per_thread
export NDT_B200_NO_GRAPH=1
for w in config4; do
python tools/perf_frame.py $w 1 > gpurun_out/plain_$w.log 2>&1 && ncu --metrics $M --clock-control none --csv --log-file gpurun_out/r02_smem_$w.csv python tools/perf_frame.py $w 1 > gpurun_out/ncu_$w.log 2>&1
done
cat $L
