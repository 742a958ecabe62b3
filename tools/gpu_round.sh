#!/bin/bash
# One GPU session: parity tests, bench line, ncu launch list of the bench command, one full ncu capture of the
# dominant kernel.  Run as:  gpurun --timeout 1700 -- 'bash tools/gpu_round.sh'
# Outputs land in gpurun_out/ (scratch); summaries worth keeping are copied to profiles/ afterwards.
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.max.mem,power.limit --format=csv > gpurun_out/gpu.txt 2>&1
STEP=${1:-all}
if [ "$STEP" = all ] || [ "$STEP" = tests ]; then
  timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1
  echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
  tail -5 gpurun_out/pytest_gpu.log
fi
if [ "$STEP" = all ] || [ "$STEP" = bench ]; then
  timeout 600 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err
  echo "bench exit $?"; cat gpurun_out/bench.json
  timeout 600 python bench.py --impl reference > gpurun_out/bench_ref.json 2>> gpurun_out/bench.err
  cat gpurun_out/bench_ref.json
fi
if [ "$STEP" = all ] || [ "$STEP" = ncu ]; then
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches.csv \
      python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e --no-xs > gpurun_out/bench_under_ncu.log 2>&1
  python tools/ncu_summary.py launches gpurun_out/launches.csv > gpurun_out/launches_summary.txt 2>&1
  cat gpurun_out/launches_summary.txt
  timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:${KERNEL:-k_walk} -c 1 \
      -f -o gpurun_out/${KERNEL:-k_walk} python tools/prof_driver.py --samples 1e7 --cycles 3 --profile-cycle 2 > gpurun_out/ncu_full.log 2>&1
  ncu -i gpurun_out/${KERNEL:-k_walk}.ncu-rep --page raw --csv > gpurun_out/${KERNEL:-k_walk}_raw.csv 2>/dev/null
  python tools/ncu_summary.py raw gpurun_out/${KERNEL:-k_walk}_raw.csv > gpurun_out/${KERNEL:-k_walk}_summary.txt 2>&1
  cat gpurun_out/${KERNEL:-k_walk}_summary.txt
fi
