#!/bin/bash
# compute-sanitizer over the paths added late in round 2 (pytest subsets through the C ABI)
set -u
O=gpurun_out
K='stphy_ac or ragged or EPI_ or predrnn_ln_3x32 or trajgru_3x32'
for t in memcheck racecheck synccheck; do
  timeout 900 compute-sanitizer --tool $t python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "$K" > $O/r02_san2_$t.log 2>&1
  echo "$t: $(grep -E 'passed|failed' $O/r02_san2_$t.log | tail -1) | $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' $O/r02_san2_$t.log | tail -1)"
done
