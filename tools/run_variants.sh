export PEMAP_BENCH_GENOME=1000000000 PEMAP_BENCH_PAIRS=2097152
for v in v1 v3; do
  PEMAP_LIB=$PWD/pecaller_b200/libpemap_$v.so python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_$v.json 2> gpurun_out/bench_$v.err
done
python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_v0.json 2> gpurun_out/bench_v0.err
