#!/bin/bash
# build a tuning variant of libmcb200.so: tools/build_variant.sh <name> <extra nvcc flags...>   (output: mc_old_b200/variants/<name>.so)
set -e
name=$1; shift
cd "$(dirname "$0")/.."
mkdir -p mc_old_b200/variants /tmp/mcbv_$name
F="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -fmad=false -Xcompiler -fPIC -Iinclude -Imc_old_b200/csrc"
nvcc $F "$@" -c mc_old_b200/csrc/mcb_kernels.cu -o /tmp/mcbv_$name/k.o &
nvcc $F "$@" -c mc_old_b200/csrc/mcb_api.cu -o /tmp/mcbv_$name/a.o &
nvcc $F "$@" -c mc_old_b200/csrc/mcb_walk.cu -o /tmp/mcbv_$name/w.o &
wait
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o mc_old_b200/variants/$name.so /tmp/mcbv_$name/k.o /tmp/mcbv_$name/a.o /tmp/mcbv_$name/w.o mc_old_b200/build/mcb_tables.o -lcudart -ldl
echo built $name
