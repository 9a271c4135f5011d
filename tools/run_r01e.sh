# deferred band pass + cooperative walk: parity (both DP variants), then cfg2 bench for both
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_e1.log 2>&1; echo "rc_defer=$?" >> gpurun_out/pytest_e1.log
PEMAP_TRACE_DEFER=0 python -m pytest tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/pytest_e0.log 2>&1; echo "rc_nodefer=$?" >> gpurun_out/pytest_e0.log
PEMAP_VERBOSE=1 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_e1.json 2> gpurun_out/bench_e1.err
PEMAP_VERBOSE=1 PEMAP_TRACE_DEFER=0 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_e0.json 2> gpurun_out/bench_e0.err
tail -n 3 gpurun_out/pytest_e1.log gpurun_out/pytest_e0.log
