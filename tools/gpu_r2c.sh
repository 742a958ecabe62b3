#!/bin/bash
set -u
mkdir -p gpurun_out; rm -f gpurun_out/sweep.txt
timeout 900 python -m pytest tests -m gpu -q --timeout=300 -x -k "functions or division or heu_history or tallies_history or event_queue" > gpurun_out/pytest_gpu3.log 2>&1
echo "pytest exit $?"; grep -v "^$" gpurun_out/pytest_gpu3.log | tail -8
bash tools/sweep.sh "base ni gs nigs nigs5 gs5 gs4" "X=1"
bash tools/sweep.sh "base gs" "MCB_HASH_BITS=12 MCB_HASH_BITS=16"
timeout 300 python bench.py --no-cpu --no-e2e --steps 5 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(json.dumps(d['xs_lookup_microbench']))"
