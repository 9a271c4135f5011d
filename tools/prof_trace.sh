# one full capture of the integer traceback kernel (third launch), small bench
export PEMAP_BENCH_PAIRS=1048576
CMD="python bench.py --steps 1 --warmup 3 --no-cpu-baseline"
ncu --set full --clock-control none --import-source on -k "regex:k_trace_i16" -s 3 -c 1 -f -o gpurun_out/prof_trace_e $CMD > gpurun_out/ncu_trace_e.log 2>&1
tail -n 3 gpurun_out/ncu_trace_e.log
