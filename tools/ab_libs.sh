#!/bin/bash
# A/B/C... of library variants on one box, interleaved:  bash tools/ab_libs.sh [-r reps] <name|base> ...   (GPU box only)
reps=3
if [ "$1" = "-r" ]; then reps=$2; shift; shift; fi
for rep in $(seq $reps); do
  for v in "$@"; do
    lib=tochris_b200/libref_cuda_$v.so; [ "$v" = base ] && lib=tochris_b200/libref_cuda.so
    RC_LIBPATH=$lib python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/ab_tmp.json 2>/dev/null
    python -c "
import json; d=json.load(open('gpurun_out/ab_tmp.json')); print('[$v]', d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], d['value_verified']['fb_sha_matches_golden'], d['status'])"
  done
done
rm -f gpurun_out/ab_tmp.json
