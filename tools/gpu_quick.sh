#!/bin/bash
# quick GPU iteration: the C2 bench line only (no extras, no CPU baseline), per-kernel times printed
tag=${1:-q}; shift
/usr/local/graft/bin/gpurun --timeout 600 -- "timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-extras $* > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; python -c \"import json; d=json.load(open('gpurun_out/${tag}_bench.json')); print(d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], d['kernel_ms_per_step'], d['status'], d['value_verified'])\"; tail -3 gpurun_out/${tag}_bench.err" 2>&1 | tail -5
