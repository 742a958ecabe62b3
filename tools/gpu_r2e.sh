#!/bin/bash
# tally scoring: parity + rates
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout=300 -x -k "tallies or trmm or output or walk_forms or fixed_source or leak" > gpurun_out/pytest_r2e.log 2>&1
echo "pytest exit $?"; grep -v "^$" gpurun_out/pytest_r2e.log | tail -6
timeout 600 python tools/deck_rates.py 2>&1 | tee gpurun_out/rates.txt
timeout 200 python tools/trmm_rate.py 2>&1 | tail -1
