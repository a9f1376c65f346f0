#!/bin/bash
# phase cycle counts of k_front (debug build with -DRC_FRONT_TIMING), then restores the normal build
cd "$(dirname "$0")/../tochris_b200/csrc" && touch rc_gpu.cu && make -s NVCCFLAGS='-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo --fmad=false -std=c++17 -Xcompiler -fPIC,-ffp-contract=off,-fno-strict-aliasing -Xptxas -v -DRC_FRONT_TIMING' 2>&1 | grep -i "error"
cd ../..
/usr/local/graft/bin/gpurun --timeout 600 -- 'timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline 2>&1 | grep k_front | tail -2' 2>&1 | tail -4
cd tochris_b200/csrc && touch rc_gpu.cu && make -s 2>&1 | grep -i error
