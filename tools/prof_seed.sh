export PEMAP_BENCH_PAIRS=524288
CMD="python bench.py --steps 1 --warmup 3 --no-cpu-baseline"
PEMAP_VERBOSE=1 $CMD > gpurun_out/plain_seed.json 2> gpurun_out/plain_seed.err && \
ncu --set full --clock-control none --import-source on -k regex:'k_seed_chain' -s 13 -c 2 -o gpurun_out/prof_seed $CMD > gpurun_out/ncu_seed.log 2>&1
tail -2 gpurun_out/ncu_seed.log; cat gpurun_out/plain_seed.err | tail -2
python - <<PY
import json
d=json.load(open("gpurun_out/plain_seed.json"))
print(d["value"], d["stage_ms_per_step"])
PY
