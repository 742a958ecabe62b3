#!/bin/bash
# round 2f, final state: the whole GPU suite and smoke() as the driver runs them, the bench line, the example decks at
# their shipped sizes next to the reference, launch list of the bench command, fresh full captures of the walk kernel
# (HEU, non-scoring) and of the xs_lookup kernel
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout=400 > gpurun_out/pytest_${TAG:-r2f}.log 2>&1
echo "pytest exit $?"; grep -v "^$" gpurun_out/pytest_${TAG:-r2f}.log | tail -4
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 600 python bench.py > gpurun_out/${TAG:-r2f}_bench.json 2> gpurun_out/${TAG:-r2f}_bench.err
echo "bench exit $?"; cut -c1-300 gpurun_out/${TAG:-r2f}_bench.json
[ -n "${SKIP_SHIPPED:-}" ] || timeout 600 python tools/shipped_decks.py --ref-timeout 75 > gpurun_out/${TAG:-r2f}_shipped_decks.jsonl 2> gpurun_out/${TAG:-r2f}_shipped_decks.err
echo "shipped exit $?"; cut -c1-420 gpurun_out/${TAG:-r2f}_shipped_decks.jsonl | head -9
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/${TAG:-r2f}_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e --no-xs > gpurun_out/bench_under_ncu.log 2>&1
python tools/ncu_summary.py launches gpurun_out/${TAG:-r2f}_launches.csv > gpurun_out/${TAG:-r2f}_launches_bench_steps2.txt 2>&1
cat gpurun_out/${TAG:-r2f}_launches_bench_steps2.txt
bash tools/gpu_prof.sh ${TAG:-r2f}_k_walk k_walk --deck heu --samples 1e7 | head -30
python tools/ncu_lowlane.py gpurun_out/${TAG:-r2f}_k_walk_src.csv > gpurun_out/${TAG:-r2f}_k_walk_per_instruction.txt 2>&1
timeout 300 ncu --set full --clock-control none -k regex:k_xs_lookup -c 2 -f -o gpurun_out/${TAG:-r2f}_xs_lookup \
    python bench.py --steps 1 --warmup 3 --no-cpu --no-e2e > gpurun_out/ncu_xs.log 2>&1
ncu -i gpurun_out/${TAG:-r2f}_xs_lookup.ncu-rep --page raw --csv > gpurun_out/${TAG:-r2f}_xs_lookup_raw.csv 2>/dev/null
python tools/ncu_summary.py raw gpurun_out/${TAG:-r2f}_xs_lookup_raw.csv > gpurun_out/${TAG:-r2f}_xs_lookup_summary.txt 2>&1
cat gpurun_out/${TAG:-r2f}_xs_lookup_summary.txt
rm -f gpurun_out/${TAG:-r2f}_xs_lookup.ncu-rep
