#!/bin/bash
# build a tuning variant of libmcb200.so: tools/build_variant.sh <name> <extra nvcc flags...>   (output: mc_old_b200/variants/<name>.so)
set -e
name=$1; shift
cd "$(dirname "$0")/.."
mkdir -p mc_old_b200/variants /tmp/mcbv_$name
F="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -fmad=false -Xcompiler -fPIC -Iinclude -Imc_old_b200/csrc"
objs=""
for f in mc_old_b200/csrc/*.cu; do
  o=/tmp/mcbv_$name/$(basename $f .cu).o
  nvcc $F "$@" -c $f -o $o &
  objs="$objs $o"
done
wait
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o mc_old_b200/variants/$name.so $objs mc_old_b200/build/mcb_tables.o -lcudart -ldl
echo built $name
