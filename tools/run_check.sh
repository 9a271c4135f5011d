python -m pytest tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/pytest_t.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_t.log
tail -n 4 gpurun_out/pytest_t.log
python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_t.json 2> gpurun_out/bench_t.err
