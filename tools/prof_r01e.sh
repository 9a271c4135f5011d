# Round-1 final profile (run under gpurun): plain run, launch list, one full capture per kernel.
export PEMAP_BENCH_PAIRS=1048576
CMD="python bench.py --steps 1 --warmup 3 --no-cpu-baseline"
K="regex:k_seed_chain|k_diag_certify|k_sw_i16|k_select|k_apply_diag|k_trace_dp16|k_trace_walk16|k_sw_fp64"
$CMD > gpurun_out/plain_r01e.json 2> gpurun_out/plain_r01e.err && \
ncu --metrics gpu__time_duration.sum --clock-control none -k "$K" -c 400 --csv --log-file gpurun_out/launches_r01e.csv $CMD > gpurun_out/ncu_launch_r01e.log 2>&1 ; \
ncu --set full --clock-control none --import-source on -k "$K" -s 66 -c 11 -f -o gpurun_out/prof_r01e $CMD > gpurun_out/ncu_full_r01e.log 2>&1
tail -n 3 gpurun_out/ncu_full_r01e.log
