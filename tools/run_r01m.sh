python -m pytest tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/pytest_n.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_n.log
tail -n 4 gpurun_out/pytest_n.log
PEMAP_VERBOSE=1 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_n.json 2> gpurun_out/bench_n.err
PEMAP_LIB=$PWD/pecaller_b200/libpemap_dbg.so PEMAP_BENCH_PAIRS=2097152 python bench.py --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/bench_dbg.json 2> gpurun_out/bench_dbg.err; grep "tie debug" gpurun_out/bench_dbg.err
