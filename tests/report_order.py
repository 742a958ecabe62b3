"""Maps the reference's output tree (/<estimator>/<score>/{mean,uncertainty}, Estimator.cpp:368-422) onto the flat
tally vector of the C-ABI (mcb_get_tallies: [estimator][score][filter bins], Estimator.cpp:280-295)."""
import numpy as np


def flatten(deck, rec):
    nt = deck.info["n_tallies"]
    mean = np.zeros(nt); uncer = np.zeros(nt)
    for est in deck.estimators():
        bins = est["n_tallies"] // max(len(est["scores"]), 1)
        for k, score in enumerate(est["scores"]):
            b = est["tally_begin"] + k * bins
            mean[b:b + bins] = rec["/%s/%s/mean" % (est["name"], score)]
            uncer[b:b + bins] = rec["/%s/%s/uncertainty" % (est["name"], score)]
    return mean, uncer
