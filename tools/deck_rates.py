#!/usr/bin/env python3
"""Histories/s of the other BASELINE configs on one GPU (not the bench metric; orientation for tuning)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mc_old_b200 as mcb
from mc_old_b200 import decks
cases = [("slab_analytic", decks.slab(samples=10_000_000), 2), ("shielding_vReduction", decks.shielding(samples=10_000_000), 2),
         ("infinite_GCR (k only)", decks.gcr(samples=200_000, active=0, passive=3), 3),
         ("infinite_GCR_TRMM", decks.gcr(samples=200_000, active=2, passive=1, trmm=True), 3),
         ("UCube", decks.ucube(samples=2_000_000, active=2, passive=2), 4),
         ("HEU_sphere_leakage", decks.heu_leakage(samples=4_000_000), 2)]
for name, xml, cycles in cases:
    deck = mcb.Deck(xml=xml)
    if not deck.info["ksearch"]:
        deck.set_run(n_cycle=cycles, n_passive=0)
    ctx = mcb.Context(deck)
    rs = []
    for c in range(cycles):
        t = time.perf_counter(); r = ctx.run_cycle(); rs.append((time.perf_counter() - t, r))
    dt, r = rs[-1]
    print("%-24s %9d hist/cycle  last cycle %.1f ms  %.3e hist/s  %.3e tracks/s  (%.1f tracks/hist)" %
          (name, r.n_histories, dt * 1e3, r.n_histories / dt, r.n_tracks / dt, r.n_tracks / r.n_histories), flush=True)
    ctx.close()
