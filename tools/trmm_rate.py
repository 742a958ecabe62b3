#!/usr/bin/env python3
"""Histories/s of the infinite_GCR_TRMM deck (BASELINE configs[2]) on one GPU; orientation for tuning the scoring path."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mc_old_b200 as mcb
from mc_old_b200 import decks
n = int(sys.argv[1]) if len(sys.argv) > 1 else 200_000
ctx = mcb.Context(mcb.Deck(xml=decks.gcr(samples=n, active=3, passive=1, trmm=True)))
for c in range(4):
    t = time.perf_counter(); r = ctx.run_cycle(); dt = time.perf_counter() - t
print("%s TRMM %d hist/cycle  last cycle %.1f ms  %.3e hist/s  k=%.6f" % (os.environ.get("MCB200_LIB", "default").split("/")[-1], n, dt * 1e3, r.n_histories / dt, r.k_cycle), flush=True)
