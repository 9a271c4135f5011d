export PEMAP_BENCH_PAIRS=4194304
for v in v1 v2; do
  PEMAP_LIB=$PWD/pecaller_b200/libpemap_$v.so python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/bench_$v.json 2> gpurun_out/bench_$v.err
done
PEMAP_LIB=$PWD/pecaller_b200/libpemap_v2.so PEMAP_FILTER_WINDOW=0 python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/bench_v3.json 2> gpurun_out/bench_v3.err
python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/bench_v0.json 2> gpurun_out/bench_v0.err
