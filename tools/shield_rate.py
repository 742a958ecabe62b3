#!/usr/bin/env python3
"""shielding_vReduction alone, 1e7 histories per cycle, a few cycles: ms per cycle (tuning aid; env knobs pass through)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mc_old_b200 as mcb
from mc_old_b200 import decks
deck = mcb.Deck(xml=decks.shielding(samples=10_000_000))
deck.set_run(n_cycle=4, n_passive=0)
ctx = mcb.Context(deck)
ts = []
for c in range(4):
    t = time.perf_counter(); r = ctx.run_cycle(); ts.append((time.perf_counter() - t) * 1e3)
print("shielding 1e7: cycles %s ms  tracks/hist %.2f" % (" ".join("%.1f" % t for t in ts), r.n_tracks / r.n_histories))
ctx.close()
