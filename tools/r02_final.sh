#!/bin/bash
# Final round-2 numbers: bench lines (both arms) and per-layer profiles of every workload.
set -u
O=gpurun_out
python bench.py > $O/r02_bench_final2.json 2> $O/r02_bench_final2.err
python bench.py --impl reference > $O/r02_ref_final2.json 2> $O/r02_ref_final2.err
for w in "cfg5 171" "cfg4 256" "cfg2 256" "cfg3 256" "cfg3ln 256"; do
  set -- $w
  python tools/layer_profile.py $1 $2 > $O/r02_layers_final_$1.txt 2>&1
done
tail -c 400 $O/r02_bench_final2.json
