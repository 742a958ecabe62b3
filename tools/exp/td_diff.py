import sys, os
R = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, R); sys.path.insert(0, os.path.join(R, "tests"))
import numpy as np
import mc_old_b200 as mcb
from mc_old_b200 import decks
import oracle_lib as ol
def run(n):
    deck = mcb.Deck(xml=decks.gcr_td(samples=n))
    ctx = mcb.Context(deck, device=0)
    orc = ol.Oracle(deck, rng_mode=ol.RNG_HISTORY, pick_mode=ol.PICK_FLOOR)
    g = ctx.run_cycle(); o = orc.run_cycle()
    orc.end_simulation()
    gm, gu = ctx.tallies(); om, ou = orc.tallies()
    ctx.close()
    return gm * n, om * n, g, o
def differs(n):
    gs, os_, g, o = run(n)
    return np.any(np.abs(gs - os_) > 1e-9 * np.abs(os_))
lo, hi = 0, int(sys.argv[1]) if len(sys.argv) > 1 else 3000
assert differs(hi)
while hi - lo > 1:
    mid = (lo + hi) // 2
    if differs(mid): hi = mid
    else: lo = mid
print("first differing history index", hi - 1)
g1, o1, g, o = run(hi); g0, o0, _, _ = run(hi - 1) if hi > 1 else (0 * g1, 0 * o1, None, None)
dg, do = g1 - g0, o1 - o0
print("tracks", g.n_tracks, o.n_tracks)
for b in np.where((np.abs(dg) > 0) | (np.abs(do) > 0))[0]:
    print("bin", b, "census", b // 50, "group", b % 50, "gpu %.9g oracle %.9g" % (dg[b], do[b]), "" if abs(dg[b] - do[b]) <= 1e-7 * abs(do[b]) else "  <<<")
