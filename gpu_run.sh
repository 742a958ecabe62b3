python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
for M in 2 3; do echo "== fused M=$M"; MCB_STEP_EVENTS=$M python tools/prof_driver.py --samples 1e7 --cycles 4 | tail -1; done
echo "== split"; MCB_MODE=split python tools/prof_driver.py --samples 1e7 --cycles 4 | tail -1
