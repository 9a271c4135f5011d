# Round-1 profile of the committed path (run under gpurun): plain run, launch list, one full capture per kernel.
export PEMAP_BENCH_PAIRS=1048576
CMD="python bench.py --steps 1 --warmup 3 --no-cpu-baseline"
K="regex:k_seed_chain|k_sw_i16|k_select|k_apply_diag|k_trace_i16|k_sw_fp64|k_len_range"
$CMD > gpurun_out/plain_r01d.json 2> gpurun_out/plain_r01d.err && \
ncu --metrics gpu__time_duration.sum --clock-control none -k "$K" -s 225 -c 80 --csv --log-file gpurun_out/launches_r01d.csv $CMD > gpurun_out/ncu_launch_r01d.log 2>&1 ; \
ncu --set full --clock-control none --import-source on -k "regex:k_seed_chain|k_sw_i16|k_apply_diag|k_trace_i16|k_sw_fp64" -s 207 -c 9 -o gpurun_out/prof_r01d $CMD > gpurun_out/ncu_full_r01d.log 2>&1
tail -n 3 gpurun_out/ncu_full_r01d.log
