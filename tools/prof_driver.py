#!/usr/bin/env python3
"""Profiling driver: a few generations of the HEU sphere through the C-ABI (to be wrapped by ncu, see profiles/README.md)."""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mc_old_b200 as mcb
from mc_old_b200 import decks

ap = argparse.ArgumentParser()
ap.add_argument("--samples", type=float, default=1e7)
ap.add_argument("--cycles", type=int, default=3)
ap.add_argument("--deck", default="heu")
ap.add_argument("--profile-cycle", type=int, default=-1, help="cudaProfilerStart() before this cycle, Stop after it (ncu --profile-from-start off)")
a = ap.parse_args()
xml = {"heu": lambda: decks.heu_sphere(samples=int(a.samples), active=a.cycles, passive=0),
       "ucube": lambda: decks.ucube(samples=int(a.samples), active=a.cycles, passive=0),
       "gcr": lambda: decks.gcr(samples=int(a.samples), active=a.cycles, passive=0),
       "gcr_trmm": lambda: decks.gcr(samples=int(a.samples), active=a.cycles, passive=0, trmm=True),
       "shield": lambda: decks.shielding(samples=int(a.samples)),
       "leak": lambda: decks.heu_leakage(samples=int(a.samples))}[a.deck]()
deck = mcb.Deck(xml=xml)
if not deck.info["ksearch"]:
    deck.set_run(n_cycle=a.cycles, n_passive=0)
ctx = mcb.Context(deck)
import ctypes
rt = ctypes.CDLL("libcudart.so")
for c in range(a.cycles):
    if c == a.profile_cycle:
        rt.cudaProfilerStart()
    r = ctx.run_cycle()
    if c == a.profile_cycle:
        rt.cudaProfilerStop()
    print(c, r.k_cycle, r.n_tracks, r.n_collisions, r.n_iterations, r.ms_transport, flush=True)
ctx.close()
