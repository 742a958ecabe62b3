#!/bin/bash
# tuning sweep on the GPU box: every variant library (mc_old_b200/variants/<name>.so) x every "ENV=VAL" setting
# usage: tools/sweep.sh "<variant names>" "<env settings, e.g. MCB_HASH_BITS=12 MCB_HASH_BITS=14>" [extra bench args]
mkdir -p gpurun_out
for v in $1; do for ev in $2; do
  r=$(env $ev MCB200_LIB=$PWD/mc_old_b200/variants/$v.so timeout 200 python bench.py --no-cpu --no-e2e --no-xs --steps 10 --warmup 3 $3 2>>gpurun_out/sweep.err | python -c "import sys,json; d=json.loads(sys.stdin.read()); s=d['stages_ms_2_generations']; print('%.3f ms/step  %.1f Mhist/s  walk %.3f ms/gen  bank %.3f source %.3f k=%.5f' % (d['ms_per_step'], d['value']/1e6, s['step']['ms']/2, s['bank']['ms']/2, s['source']['ms']/2, d['k_cycle_last']))")
  echo "$v $ev : $r" | tee -a gpurun_out/sweep.txt
done; done
