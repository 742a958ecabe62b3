#!/bin/bash
# ncu --set full captures of the per-event-type kernels (event-queue mode) and of the xs_lookup microbench kernel:
# the per-stage evidence the north star asks for (L2 hit rate / GB/s of the XS gathers, divergence and occupancy of
# the collision kernel).  Run as: gpurun --timeout 900 -- 'bash tools/ncu_stages.sh'
set -u
mkdir -p gpurun_out
for k in k_xs_stage k_flight k_collide k_cross; do
  MCB_MODE=split timeout 300 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:$k -c 1 \
      -f -o gpurun_out/split_$k python tools/prof_driver.py --samples 1e7 --cycles 3 --profile-cycle 2 > gpurun_out/ncu_$k.log 2>&1
  ncu -i gpurun_out/split_$k.ncu-rep --page raw --csv > gpurun_out/split_${k}_raw.csv 2>/dev/null
  python tools/ncu_summary.py raw gpurun_out/split_${k}_raw.csv
done > gpurun_out/split_stage_summary.txt 2>&1
cat gpurun_out/split_stage_summary.txt
timeout 300 ncu --set full --clock-control none -k regex:k_xs_lookup -c 2 -f -o gpurun_out/xs_lookup \
    python bench.py --steps 1 --warmup 3 --no-cpu --no-e2e > gpurun_out/ncu_xs.log 2>&1
ncu -i gpurun_out/xs_lookup.ncu-rep --page raw --csv > gpurun_out/xs_lookup_raw.csv 2>/dev/null
python tools/ncu_summary.py raw gpurun_out/xs_lookup_raw.csv > gpurun_out/xs_lookup_summary.txt 2>&1
cat gpurun_out/xs_lookup_summary.txt
