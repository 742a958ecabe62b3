#!/bin/bash
# round 2f: shipped-size decks in wall seconds next to the reference; the full bench line
set -u
mkdir -p gpurun_out
timeout 600 python tools/shipped_decks.py --ref-timeout 75 > gpurun_out/shipped_decks.jsonl 2> gpurun_out/shipped_decks.err
echo "shipped exit $?"; cut -c1-260 gpurun_out/shipped_decks.jsonl | head -12
timeout 600 python bench.py > gpurun_out/bench_r2f.json 2> gpurun_out/bench_r2f.err
echo "bench exit $?"; cut -c1-400 gpurun_out/bench_r2f.json
