export PEMAP_BENCH_PAIRS=524288
CMD="python bench.py --steps 1 --warmup 3 --no-cpu-baseline"
$CMD > gpurun_out/plain_trace.json 2> gpurun_out/plain_trace.err && \
ncu --set full --clock-control none --import-source on -k regex:'k_trace_i16|k_sw_fp64' -s 24 -c 4 -o gpurun_out/prof_trace $CMD > gpurun_out/ncu_trace.log 2>&1
tail -n 2 gpurun_out/ncu_trace.log
