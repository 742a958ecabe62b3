#!/bin/bash
# round 2f: <disk_z> source tests, xs_lookup variants, walk-kernel variants
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q --timeout=300 -x -k "sphere_det or disk or slab" > gpurun_out/pytest_r2f.log 2>&1
echo "pytest exit $?"; grep -v "^$" gpurun_out/pytest_r2f.log | tail -6
for v in "" xsl_p1 xsl_p2_b2 xsl_p2_b4; do
  if [ -z "$v" ]; then timeout 120 python tools/xs_rate.py 2>&1 | tail -1; else MCB200_LIB=$PWD/mc_old_b200/variants/$v.so timeout 120 python tools/xs_rate.py 2>&1 | tail -1; fi
done | tee gpurun_out/xs_rates_r2f.txt
rm -f gpurun_out/sweep.txt
cp mc_old_b200/libmcb200.so mc_old_b200/variants/main.so
tools/sweep.sh "main walk_p2 minb5 minb6" "X=0" > /dev/null
cat gpurun_out/sweep.txt
