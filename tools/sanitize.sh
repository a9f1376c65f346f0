#!/bin/bash
# compute-sanitizer over the hot path (the reference's own race / memory tooling is none; ours is this).
# GPU box only:  gpurun --timeout 1500 -- 'bash tools/sanitize.sh r02a'
# Writes gpurun_out/<tag>_sanitize_<tool>_<case>.log; copy the summaries worth keeping into profiles/.
tag=${1:-r02}
out=gpurun_out
mkdir -p $out
CS=/usr/local/cuda/bin/compute-sanitizer
world="--scene small --w 320 --h 200 --n 8 --async-calls 4"
ents="--scene full --w 320 --h 240 --n 4 --entities 6 --particles 100 --sprites 3 --doors 2 --dlights 2 --underwater_every 3"
for tool in memcheck racecheck synccheck initcheck; do
  for c in world ents; do
    args=${!c}
    extra=""
    [ $tool = racecheck ] && extra="--racecheck-report all"
    [ $tool = initcheck ] && extra="--track-unused-memory no"
    timeout 600 $CS --tool $tool $extra --print-limit 20 --log-file $out/${tag}_sanitize_${tool}_${c}.log \
        python tools/render_case.py $args --no-oracle > $out/${tag}_sanitize_${tool}_${c}.out 2>&1
    echo "$tool $c exit $?" >> $out/${tag}_sanitize_summary.txt
    tail -3 $out/${tag}_sanitize_${tool}_${c}.log >> $out/${tag}_sanitize_summary.txt
  done
done
cat $out/${tag}_sanitize_summary.txt
