#!/bin/bash
# Round-2 measurement set on one B200 (run under gpurun from the repo root); everything lands in gpurun_out/ and the
# summaries are copied to profiles/ by hand (profiles/README_r02.md says which file is which).
# Plain runs first (each must exit 0), the ncu captures of the same commands afterwards.
set -u
O=gpurun_out
mkdir -p $O
run() { echo "== $*" >&2; "$@"; echo "== rc=$?" >&2; }
run ./build_tools/chunk_probe 64 > $O/chunk_probe_64gb_r02.json 2> $O/chunk_probe_64.err
run ./build_tools/alu_peak > $O/alu_peak_raw_r02.json 2> $O/alu_peak.err
run python tools/cfg3_check.py 4000000 > $O/cfg3_check_r02.json 2> $O/cfg3_check.err
run python tools/cfg5_check.py > $O/cfg5_check_r02.json 2> $O/cfg5_check.err
run python tools/cfg4_sweep.py > $O/cfg4_sweep_r02.json 2> $O/cfg4_sweep.err
run python tools/sw_sweep.py > $O/sw_sweep_r02.json 2> $O/sw_sweep.err
run python tools/cfg3_parity.py 2000000 > $O/cfg3_parity_r02.json 2> $O/cfg3_parity.err
PEMAP_PARITY_CONFIG=cfg5 PEMAP_PARITY_ALSO_SINGLE=400000 run python tools/cfg3_parity.py 200000 > $O/cfg5_parity_r02.json 2> $O/cfg5_parity.err
run python tools/cli_e2e.py 10000000 > $O/cli_e2e_r02.json 2> $O/cli_e2e.err
run python bench.py --steps 5 --warmup 3 > $O/bench_cfg3_r02.json 2> $O/bench_cfg3.err
run python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_ref_cfg3_r02.json 2> $O/bench_ref_cfg3.err
run python bench.py --config cfg2 --steps 5 --warmup 3 > $O/bench_cfg2_r02.json 2> $O/bench_cfg2.err
# ncu: launch list of every kernel of ours over four passes of 1 M pairs on cfg3 (the last pass is the warm one), then
# --set full of the dominant kernels, one launch each (one lane = two chunks of 524288 pairs per pass: k_seed_rbi matches three
# launches per chunk, the DP regex four; the warm pass starts at chunk 6)
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_ -c 900 --csv --log-file $O/launches_r02.csv \
    python tools/cfg3_check.py 1048576 > $O/ncu_launches.log 2>&1
ncu --set full --import-source on --clock-control none --kernel-name regex:k_seed_rbi --launch-skip 18 --launch-count 1 \
    -o $O/ncu_seed_rbi_r02 python tools/cfg3_check.py 1048576 > $O/ncu_seed.log 2>&1
ncu --set full --clock-control none --kernel-name regex:"k_sw_i16|k_trace_dp16|k_trace_walk16|k_diag_certify" --launch-skip 24 --launch-count 4 \
    -o $O/ncu_dp_r02 python tools/cfg3_check.py 1048576 > $O/ncu_dp.log 2>&1
ls -la $O/*r02* >&2
