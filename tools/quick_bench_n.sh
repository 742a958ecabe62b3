#!/bin/bash
# quick_bench for N ranks under torchrun: quick_bench_n.sh N "ENV=a" "ENV=b" ...
N=$1; shift
for ev in "$@"; do
  r=$(env $ev timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --no-cpu --no-xs --steps 10 --warmup 3 $BENCH_ARGS 2>>gpurun_out/quick.err | grep '^{' | python -c "import sys,json; d=json.loads(sys.stdin.read()); s=d['stages_ms_2_generations']; e=d.get('e2e') or {}; print('%.3f ms/step %.1f Mhist/s walk %.3f source %.3f bank %.3f closeout %.3f e2e %.2f ms k=%.5f' % (d['ms_per_step'], d['value']/1e6, s['step']['ms']/2, s['source']['ms']/2, s['bank']['ms']/2, s['closeout']['ms']/2, e.get('ms_per_step', 0), d['k_cycle_last']))")
  echo "N=$N $ev : $r"
done
