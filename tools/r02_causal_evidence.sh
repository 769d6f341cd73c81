#!/bin/bash
# Evidence run for the Causal LSTM + GHU path (one gpurun call): full GPU suite, bench line, ncu launch list + one full
# capture of a cfg3pp step, sanitizer passes over the new tests.
set -u
O=gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > $O/s5_tests.log 2>&1
tail -2 $O/s5_tests.log
timeout 600 python bench.py > $O/s5_bench.json 2> $O/s5_bench.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/r02_launches_cfg3pp_b64.csv \
    python tools/run_once.py cfg3pp 64 1 > $O/s5_launches.log 2>&1
# one whole timestep of the second prediction step: 4 cells x (C, M, O.conv_last, O.conv_o) + GHU + head = 18 conv_halo launches
timeout 900 ncu --set full --clock-control none --import-source on -k regex:conv_halo_kernel --launch-skip 36 --launch-count 18 \
    -f -o $O/r02_ncu_cfg3pp_step_b256 python tools/run_once.py cfg3pp 256 1 > $O/s5_ncu.log 2>&1
for t in memcheck racecheck synccheck initcheck; do
  timeout 600 compute-sanitizer --tool $t python -m pytest tests/test_gpu_causal.py -m gpu -q -x -k "3x32 or 1x40 or single_step" > $O/r02_san_causal_$t.log 2>&1
  echo "$t: $(grep -E 'passed|failed' $O/r02_san_causal_$t.log | tail -1) | $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' $O/r02_san_causal_$t.log | tail -1)"
done
