export PEMAP_BENCH_PAIRS=1048576
CMD="python bench.py --steps 1 --warmup 3 --no-cpu-baseline"
$CMD > gpurun_out/plain_r01b.json 2> gpurun_out/plain_r01b.err && \
ncu --metrics gpu__time_duration.sum --clock-control none -s 96 -c 40 --csv --log-file gpurun_out/launches_r01b.csv $CMD > gpurun_out/ncu_launch_r01b.log 2>&1 ; \
ncu --set full --clock-control none --import-source on -k regex:'k_seed_chain|k_sw_i16|k_apply_diag|k_sw_fp64' -s 36 -c 8 -o gpurun_out/prof_r01b $CMD > gpurun_out/ncu_full_r01b.log 2>&1
tail -3 gpurun_out/ncu_full_r01b.log
