#!/bin/bash
# round 2 evidence: launch list of the bench command, full captures of the walk kernel (HEU, non-scoring), of its scoring
# instance on infinite_GCR_TRMM and on the shielding deck, and of the xs_lookup microbench kernel
set -u
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e --no-xs > gpurun_out/bench_under_ncu.log 2>&1
python tools/ncu_summary.py launches gpurun_out/launches.csv > gpurun_out/launches_summary.txt 2>&1
cat gpurun_out/launches_summary.txt
bash tools/gpu_prof.sh r2c_k_walk k_walk --deck heu --samples 1e7 | head -26
bash tools/gpu_prof.sh r2c_k_walk_trmm k_walk --deck gcr_trmm --samples 1e5 | head -26
head -24 gpurun_out/r2c_k_walk_trmm_functions.txt
bash tools/gpu_prof.sh r2c_k_walk_shield k_walk --deck shield --samples 5e6 | head -26
timeout 300 ncu --set full --clock-control none -k regex:k_xs_lookup -c 2 -f -o gpurun_out/xs_lookup \
    python bench.py --steps 1 --warmup 3 --no-cpu --no-e2e > gpurun_out/ncu_xs.log 2>&1
ncu -i gpurun_out/xs_lookup.ncu-rep --page raw --csv > gpurun_out/xs_lookup_raw.csv 2>/dev/null
python tools/ncu_summary.py raw gpurun_out/xs_lookup_raw.csv > gpurun_out/xs_lookup_summary.txt 2>&1
cat gpurun_out/xs_lookup_summary.txt
rm -f gpurun_out/xs_lookup.ncu-rep
