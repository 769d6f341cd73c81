#!/bin/bash
# Developer aid (GPU box): short A/B bench lines.   tools/ab_bench.sh <tag> [ENV=VAL ...] -- cfg1 cfg2 ...
tag=$1; shift
envs=()
while [ "$1" != "--" ] && [ $# -gt 0 ]; do envs+=("$1"); shift; done
shift
mkdir -p gpurun_out
for c in "$@"; do
  env "${envs[@]}" python bench.py --workload $c --no-cpu-baseline --steps 5 --warmup 3 > gpurun_out/ab_${tag}_$c.json 2> gpurun_out/ab_${tag}_$c.err
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/ab_${tag}_$c.json"))
    print("${tag} $c %.0f fps %.3f ms/step e2e %.0f gemm_frac %.3f" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["roofline"]["frac"]))
except Exception as e:
    print("${tag} $c FAILED", e, open("gpurun_out/ab_${tag}_$c.err").read()[-800:])
PY
done
