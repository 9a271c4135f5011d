for v in v1 v3 v6; do
  PEMAP_LIB=$PWD/pecaller_b200/libpemap_$v.so PEMAP_BENCH_PAIRS=4194304 python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/bench_$v.json 2> gpurun_out/bench_$v.err
done
PEMAP_BENCH_PAIRS=4194304 python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/bench_v0.json 2> gpurun_out/bench_v0.err
