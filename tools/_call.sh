mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 > gpurun_out/r02_pytest16.log
cat gpurun_out/r02_pytest16.log
python bench.py --steps 5 --warmup 3 > gpurun_out/r02_bench16.json 2> gpurun_out/r02_bench16.err; tail -3 gpurun_out/r02_bench16.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02_bench16_ref.json 2> gpurun_out/r02_bench16_ref.err
M=gpu__time_duration.sum,smsp__sass_thread_inst_executed_op_dadd_pred_on.sum,smsp__sass_thread_inst_executed_op_dmul_pred_on.sum,smsp__sass_thread_inst_executed_op_dfma_pred_on.sum,smsp__thread_inst_executed_pipe_fp64_pred_on.sum,smsp__inst_executed_pipe_fp64.sum,smsp__inst_executed.sum,sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_elapsed,launch__registers_per_thread,dram__bytes_read.sum,dram__bytes_write.sum
export NDT_B200_NO_GRAPH=1
for w in config2 config1 config4 config5_yaml; do
python tools/perf_frame.py $w 1 > gpurun_out/plain_$w.log 2>&1 && ncu --metrics $M --clock-control none --csv --log-file gpurun_out/r02_launches_fp64_$w.csv python tools/perf_frame.py $w 1 > gpurun_out/ncu_$w.log 2>&1
done
