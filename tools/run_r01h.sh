python -m pytest tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/pytest_h.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_h.log
tail -n 4 gpurun_out/pytest_h.log
PEMAP_VERBOSE=1 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_h.json 2> gpurun_out/bench_h.err
export PEMAP_BENCH_PAIRS=1048576
CMD="python bench.py --steps 1 --warmup 3 --no-cpu-baseline"
ncu --set full --clock-control none --import-source on -k "regex:k_trace_walk16|k_trace_dp16" -s 4 -c 2 -f -o gpurun_out/prof_trace_h $CMD > gpurun_out/ncu_trace_h.log 2>&1
tail -n 2 gpurun_out/ncu_trace_h.log
