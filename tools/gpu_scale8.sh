#!/bin/bash
# 8-GPU session: the bench line at N = 8 (1e7 histories per GPU) and the north-star upper end, 1e9 histories per generation
set -u
mkdir -p gpurun_out
run() { n=$1; shift; timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29555 bench.py --gpus $n "$@"; }
run 8 --no-cpu --no-xs --steps 10 2>gpurun_out/scale_n8.err | grep '^{' > gpurun_out/scale_n8.json
python -c "import json; d=json.load(open('gpurun_out/scale_n8.json')); print('N=8', d['ms_per_step'], d['value'], d['g_independent'], d['e2e'] and (d['e2e']['ms_per_step'], d['e2e']['value']), d['stages_ms_2_generations'])"
nvidia-smi --query-gpu=memory.used --format=csv,noheader | head -2
MCB_TRACE_MEM=1 run 8 --no-cpu --no-xs --no-e2e --samples 1.25e8 --steps 4 --warmup 3 2>gpurun_out/scale_1e9.err | grep '^{' > gpurun_out/scale_1e9.json
python -c "import json; d=json.load(open('gpurun_out/scale_1e9.json')); print('1e9/gen', d['config']['histories_per_generation'], d['ms_per_step'], d['value'], d['k_cycle_last'], d['g_independent'], d['stages_ms_2_generations'])"
grep -i "error\|mem" gpurun_out/scale_1e9.err | tail -5
