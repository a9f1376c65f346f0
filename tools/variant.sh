#!/bin/bash
# builds a kernel variant of the product library beside it:  bash tools/variant.sh <name> "<extra nvcc flags>"
#   -> tochris_b200/libref_cuda_<name>.so   (select it with RC_LIBPATH=tochris_b200/libref_cuda_<name>.so)
set -e
name=$1; flags=$2
cd "$(dirname "$0")/../tochris_b200/csrc"

make -s OUT=../libref_cuda_$name.so BUILD=../../build/var_$name EXTRA_NVCC="$flags" 2>&1 | grep -i -E "error|warning: v" || true
ls -la ../libref_cuda_$name.so
