#!/bin/bash
# full GPU check of the tree: all gpu tests, smoke, both bench arms
set -u
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q --timeout=900 > gpurun_out/pytest_full.log 2>&1
echo "pytest exit $?"; grep -v "^$" gpurun_out/pytest_full.log | tail -6
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 900 python bench.py > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err; echo "bench exit $?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_full.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','gpu_launches','g_independent') if k in d}, d['e2e'], d['roofline'], d['cpu_baseline'].get('value'))
PY
