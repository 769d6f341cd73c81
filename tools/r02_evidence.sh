#!/bin/bash
# Round-2 evidence run (one gpurun call): bench lines, ncu launch list + full capture, sanitizer passes.
set -u
O=gpurun_out
python bench.py > $O/r02_bench_final.json 2> $O/r02_bench_final.err
python bench.py --impl reference > $O/r02_ref_final.json 2> $O/r02_ref_final.err
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/r02_launches_cfg5_b32.csv \
    python bench.py --seqs-per-gpu 32 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-per-config > $O/r02_launches.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:conv_halo_kernel --launch-skip 45 --launch-count 17 \
    -f -o $O/r02_ncu_cfg5_step_b64_v2 python tools/run_once.py cfg5 64 > $O/r02_ncu_v2.log 2>&1
for t in memcheck racecheck synccheck; do
  for w in "cfg5 2" "cfg1 2" "cfg4 3"; do
    set -- $w
    timeout 600 compute-sanitizer --tool $t python tools/run_once.py $1 $2 1 > $O/r02_san_${t}_$1.log 2>&1
    echo "$t $1: $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' $O/r02_san_${t}_$1.log | tail -1)"
  done
done
tail -c 600 $O/r02_bench_final.json
