# split traceback (k_trace_dp16 + k_trace_walk16) + cooperative fp64 walk: parity, cfg2 bench, ncu capture of the new kernels
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_f.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_f.log
tail -n 4 gpurun_out/pytest_f.log
PEMAP_VERBOSE=1 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_f.json 2> gpurun_out/bench_f.err
export PEMAP_BENCH_PAIRS=1048576
CMD="python bench.py --steps 1 --warmup 3 --no-cpu-baseline"
ncu --set full --clock-control none --import-source on -k "regex:k_trace_dp16|k_trace_walk16|k_sw_fp64" -s 8 -c 4 -f -o gpurun_out/prof_trace_f $CMD > gpurun_out/ncu_trace_f.log 2>&1
tail -n 2 gpurun_out/ncu_trace_f.log
