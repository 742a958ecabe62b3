#!/bin/bash
# ncu capture of the walk kernel + the two tests that failed
set -u
mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -q --timeout=300 -k "division or host_program_cli" > gpurun_out/pytest_gpu2.log 2>&1
echo "pytest exit $?"; grep -v "^$" gpurun_out/pytest_gpu2.log | tail -15
K=${KERNEL:-k_walk}
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:$K -c 1 \
    -f -o gpurun_out/$K python tools/prof_driver.py --samples 1e7 --cycles 3 --profile-cycle 2 > gpurun_out/ncu_full.log 2>&1
ncu -i gpurun_out/$K.ncu-rep --page raw --csv > gpurun_out/${K}_raw.csv 2>/dev/null
python tools/ncu_summary.py raw gpurun_out/${K}_raw.csv > gpurun_out/${K}_summary.txt 2>&1
cat gpurun_out/${K}_summary.txt
ncu -i gpurun_out/$K.ncu-rep --page source --csv --print-source cuda,sass > gpurun_out/${K}_src.csv 2>/dev/null
python tools/ncu_regions.py gpurun_out/${K}_src.csv > gpurun_out/${K}_functions.txt 2>&1
python tools/ncu_source_lines.py gpurun_out/${K}_src.csv > gpurun_out/${K}_source_lines.txt 2>&1
head -45 gpurun_out/${K}_functions.txt
