#!/bin/bash
# Parity and timing of the phase-driven seed kernel (single-chain shortcut on / off), the diagonal fp64 scores of the
# selection and the re-framed k_sw_i16 on one B200.  The prefix tests run first: their subprocesses need the whole GPU.
set -u
O=gpurun_out
mkdir -p $O
timeout 1500 python -m pytest tests/test_gpu_cfg3_prefix.py tests/test_gpu_parity.py -x -q \
  -k "prefix or repeats_equal or layouts_agree or at_scale or cfg4 or certificate or (matches_oracle and (repeat or bis or edge9 or pe150))" \
  > $O/pytest_shortcut_r02.log 2>&1; echo "pytest rc=$?" >&2; tail -3 $O/pytest_shortcut_r02.log >&2
python tools/cfg3_check.py 4000000 > $O/cfg3_check_r02b.json 2> $O/cfg3_check_b.err; echo "cfg3 rc=$?" >&2
PEMAP_SHORTCUT=0 python tools/cfg3_check.py 4000000 > $O/cfg3_check_r02b_off.json 2> $O/cfg3_check_boff.err
python tools/cfg5_check.py > $O/cfg5_check_r02b.json 2> $O/cfg5_check_b.err; echo "cfg5 rc=$?" >&2
grep -E "ms_seed|reads_per_s|ms_select|ms_sw\"" $O/cfg3_check_r02b.json $O/cfg3_check_r02b_off.json $O/cfg5_check_r02b.json >&2
