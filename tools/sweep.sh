#!/bin/bash
# tuning sweep on the GPU box: every variant library x MCB_STEP_EVENTS; prints ms_per_step
# usage: tools/sweep.sh "<variant names>" "<events list>" [extra bench args]
mkdir -p gpurun_out
for v in $1; do for ev in $2; do
  r=$(MCB200_LIB=$PWD/mc_old_b200/variants/$v.so MCB_STEP_EVENTS=$ev timeout 200 python bench.py --no-cpu --no-e2e --steps 5 --warmup 3 $3 2>>gpurun_out/sweep.err | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('%.3f ms/step  %.1f Mhist/s  step-kernel %.3f ms/gen  bank %.3f source %.3f k=%.5f' % (d['ms_per_step'], d['value']/1e6, d['stages_ms_2_generations']['step']['ms']/2, d['stages_ms_2_generations']['bank']['ms']/2, d['stages_ms_2_generations']['source']['ms']/2, d['k_cycle_last']))")
  echo "$v events=$ev : $r" | tee -a gpurun_out/sweep.txt
done; done
