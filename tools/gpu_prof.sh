#!/bin/bash
# ncu --set full capture of one kernel through tools/prof_driver.py; summaries into gpurun_out/<tag>_*
# usage: gpu_prof.sh <tag> [kernel regex] [prof_driver args...]     (env is passed through, e.g. MCB200_LIB, MCB_WALK_EXCHANGE)
set -u
TAG=$1; K=${2:-k_walk}; shift; shift
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:$K -c 1 \
    -f -o gpurun_out/$TAG python tools/prof_driver.py --cycles 3 --profile-cycle 2 "$@" > gpurun_out/${TAG}_ncu.log 2>&1
ncu -i gpurun_out/$TAG.ncu-rep --page raw --csv > gpurun_out/${TAG}_raw.csv 2>/dev/null
python tools/ncu_summary.py raw gpurun_out/${TAG}_raw.csv > gpurun_out/${TAG}_summary.txt 2>&1
cat gpurun_out/${TAG}_summary.txt
ncu -i gpurun_out/$TAG.ncu-rep --page source --csv --print-source cuda,sass > gpurun_out/${TAG}_src.csv 2>/dev/null
python tools/ncu_regions.py gpurun_out/${TAG}_src.csv > gpurun_out/${TAG}_functions.txt 2>&1
python tools/ncu_source_lines.py gpurun_out/${TAG}_src.csv > gpurun_out/${TAG}_source_lines.txt 2>&1
python tools/ncu_stalls.py gpurun_out/${TAG}_src.csv 3 > gpurun_out/${TAG}_stalls.txt 2>&1
head -24 gpurun_out/${TAG}_stalls.txt
rm -f gpurun_out/$TAG.ncu-rep
