export PEMAP_BENCH_PAIRS=4194304 PEMAP_VERBOSE=1
for mb in 16 32; do
  PEMAP_FILTER_MB=$mb python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/bench_f$mb.json 2> gpurun_out/bench_f$mb.err
done
