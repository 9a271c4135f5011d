#!/bin/bash
# Round-2 follow-up on one B200: the whole GPU test suite, the ALU issue-rate probe with the SM clock sampled beside it,
# and the two `ncu --set full` captures (skips counted from profiles/launches_r02.csv: one lane, two chunks of 524288
# pairs per pass of 1 M pairs, four passes; k_seed_rbi matches three launches per chunk, the DP regex four).
set -u
O=gpurun_out
mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -x -q > $O/pytest_gpu_r02.log 2>&1; echo "pytest rc=$?" >&2; tail -3 $O/pytest_gpu_r02.log >&2
nvidia-smi -i 0 --query-gpu=clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active --format=csv,noheader -lms 100 > $O/alu_peak_clocks_r02.csv 2>/dev/null &
SMI=$!
./build_tools/alu_peak > $O/alu_peak_raw_r02.json 2> $O/alu_peak.err
kill $SMI
ncu --set full --import-source on --clock-control none --kernel-name regex:k_seed_rbi --launch-skip 18 --launch-count 1 \
    -o $O/ncu_seed_rbi_r02 python tools/cfg3_check.py 1048576 > $O/ncu_seed.log 2>&1
ncu --set full --clock-control none --kernel-name regex:"k_sw_i16|k_trace_dp16|k_trace_walk16|k_diag_certify" --launch-skip 24 --launch-count 4 \
    -o $O/ncu_dp_r02 python tools/cfg3_check.py 1048576 > $O/ncu_dp.log 2>&1
ls -la $O/*.ncu-rep >&2
