#!/usr/bin/env python3
"""Golden for the statistical gate of BASELINE.json config 2 (examples/infinite_GCR_TRMM): the TRMM tally set (3500 bins) and
k of a reference-identical run (oracle in RNG_GLOBAL / PICK_CDF mode, bit-identical to the compiled reference) at 4000
histories x (5 + 10) generations.  The deck's own 400 histories per generation are too few to compare a 2e4-per-generation
GPU run against: the tallies of a k-eigenvalue run carry a population-size bias of order 1 / N, visible in a chi-square
(GPU at N = 400 against the N = 400 reference: chi2 / dof 1.05 - 1.25; GPU at N = 2e4 against the same: 2.2 - 2.5).

    python tests/golden/make_gcr_trmm_golden.py      (about three minutes on one core)
"""
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

N, PASSIVE, ACTIVE = 4000, 5, 10
deck = mcb.Deck(xml=decks.gcr(samples=N, active=ACTIVE, passive=PASSIVE, trmm=True))
orc = ol.Oracle(deck, rng_mode=ol.RNG_GLOBAL, pick_mode=ol.PICK_CDF)
res = orc.run()
mean, uncer = orc.tallies()
k = np.array([r.k_cycle for r in res])
np.savez_compressed(os.path.join(HERE, "gcr_trmm_global_4000.npz"), mean=mean, uncer=uncer, k_cycle=k, k_avg=res[-1].k_avg,
                    k_uncer=res[-1].k_uncer, samples=N, passive=PASSIVE, active=ACTIVE)
print(res[-1].k_avg, res[-1].k_uncer, int((uncer > 0).sum()))
