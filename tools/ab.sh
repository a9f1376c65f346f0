#!/bin/bash
# A/B of environment-selected kernel variants on one box, interleaved (A B A B) so that box-to-box and drift effects cancel.
# usage: bash tools/ab.sh <tag> "<envA>" "<envB>" [bench args...]     (GPU box only)
tag=$1; A=$2; B=$3; shift; shift; shift
for rep in 1 2; do
  for v in A B; do
    e=${!v}
    env $e python bench.py --steps 20 --warmup 3 --no-cpu-baseline "$@" > gpurun_out/${tag}_${v}${rep}.json 2> gpurun_out/${tag}_${v}${rep}.err
    python - <<PY
import json
d=json.load(open("gpurun_out/${tag}_${v}${rep}.json"))
print("${v}${rep} [${e}]", "value", d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"]["value"], "kernels", d["kernel_ms_per_step"], "status", d["status"])
PY
  done
done
