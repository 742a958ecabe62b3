#!/usr/bin/env python3
"""Golden for the k check of the scaled-up HEU sphere (BASELINE.json config 5, north_star "k-eff within 3 sigma of the
reference"): the reference cannot run 1e8 histories per generation (~73 GB, ~500 s per generation),
so its k is pinned by an ENSEMBLE of reference-identical runs at the deck's own size: the oracle in RNG_GLOBAL / PICK_CDF
mode (bit-identical to the compiled reference, tests/test_oracle_golden.py) with 8 base seeds, 1e4 histories x (10 passive
+ 190 active) generations each.  Writes tests/golden/k_ensemble.json: k_avg and the reported k_uncer of every run.

    python tests/golden/make_k_ensemble.py        (about two minutes on one core)
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import mc_old_b200 as mcb
from mc_old_b200 import decks
import oracle_lib as ol

SEEDS = [1, 3, 5, 7, 9, 11, 13, 15]
SAMPLES, PASSIVE, ACTIVE = 10000, 10, 190
runs = []
for seed in SEEDS:
    deck = mcb.Deck(xml=decks.heu_sphere(samples=SAMPLES, active=ACTIVE, passive=PASSIVE))
    deck.set_run(seed=seed)
    res = ol.Oracle(deck, rng_mode=ol.RNG_GLOBAL, pick_mode=ol.PICK_CDF).run()
    k = np.array([r.k_cycle for r in res[PASSIVE:]])
    runs.append({"seed": seed, "k_avg": res[-1].k_avg, "k_uncer_reported": res[-1].k_uncer,
                 "k_cycle_std_over_sqrt_n": float(k.std(ddof=1) / np.sqrt(k.size))})
    print(runs[-1], flush=True)
ka = np.array([r["k_avg"] for r in runs])
out = {"deck": "HEU_sphere_criticality", "samples": SAMPLES, "passive": PASSIVE, "active": ACTIVE, "mode": "oracle RNG_GLOBAL / PICK_CDF (reference-identical)",
       "runs": runs, "ensemble_mean": float(ka.mean()), "ensemble_std_of_mean": float(ka.std(ddof=1) / np.sqrt(ka.size))}
with open(os.path.join(HERE, "k_ensemble.json"), "w") as f:
    json.dump(out, f, indent=1)
print(out["ensemble_mean"], out["ensemble_std_of_mean"])
