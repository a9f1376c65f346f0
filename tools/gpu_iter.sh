#!/bin/bash
# one GPU iteration: parity tests (optional) + bench, results under gpurun_out/<tag>_*
tag=${1:-iter}; tests=${2:-1}
cmd=""
if [ "$tests" = "1" ]; then cmd="timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_pytest_gpu.log 2>&1; tail -4 gpurun_out/${tag}_pytest_gpu.log;"; fi
cmd="$cmd timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; python -c \"import json; d=json.load(open('gpurun_out/${tag}_bench.json')); print(d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], d['kernel_ms_per_step'], d['status'])\"; tail -3 gpurun_out/${tag}_bench.err"
/usr/local/graft/bin/gpurun --timeout 1200 -- "$cmd" 2>&1 | tail -12
