#!/bin/bash
# Final check of the round on one B200: the whole GPU suite, then the default bench (cfg3) as the driver runs it.
set -u
O=gpurun_out
mkdir -p $O
timeout 1300 python -m pytest tests -m gpu -x -q > $O/pytest_gpu_final_r02.log 2>&1; echo "pytest rc=$?" >&2; tail -3 $O/pytest_gpu_final_r02.log >&2
timeout 400 python bench.py --steps 5 --warmup 3 > $O/bench_cfg3_final_r02.json 2> $O/bench_cfg3_final.err; echo "bench rc=$?" >&2
python - <<'PY' >&2
import json
d = json.loads(open("gpurun_out/bench_cfg3_final_r02.json").read().strip().splitlines()[-1])
print("value", d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"]["value"], "roof", d["roofline"]["frac"], d["roofline"]["traffic"], "cpu", d["cpu_baseline"]["value"])
PY
