#!/bin/bash
# One ncu --set full capture (with source) of one kernel on the C2 workload, plus the launch list of a whole render.
# GPU box only:  gpurun --timeout 900 -- 'bash tools/profile.sh r02e k_scan'
# -> gpurun_out/<tag>_<kernel>.ncu-rep (read here with tools/ncu_summary.py / tools/ncu_lines.py)
tag=${1:-r02}; kern=${2:-k_scan}; shift; shift
args=${@:---scene multi --w 640 --h 480 --n 64}
mkdir -p gpurun_out
python tools/render_case.py $args --reps 3 --no-oracle > gpurun_out/${tag}_${kern}_plain.log 2>&1 || { tail -5 gpurun_out/${tag}_${kern}_plain.log; exit 1; }
ncu --set full --clock-control none --import-source on -k regex:$kern --launch-skip 1 --launch-count 1 \
    -o gpurun_out/${tag}_${kern} -f python tools/render_case.py $args --reps 3 --no-oracle > gpurun_out/${tag}_${kern}_ncu.log 2>&1
tail -3 gpurun_out/${tag}_${kern}_ncu.log
