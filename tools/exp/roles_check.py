import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import mc_old_b200 as mcb
from mc_old_b200 import decks
for n in (100000, 200000, 2000000):
    out = []
    for form in ("", "roles"):
        os.environ["MCB_WALK_FORM"] = form
        deck = mcb.Deck(xml=decks.heu_sphere(samples=n, active=1, passive=2))
        ctx = mcb.Context(deck, device=0)
        rs = [ctx.run_cycle() for _ in range(2)]
        out.append([(r.k_cycle.hex(), r.n_sites, r.n_tracks, r.n_collisions, r.n_crossings) for r in rs])
        ctx.close()
    print(n, "same" if out[0] == out[1] else "DIFFERENT", out[0][1], out[1][1], flush=True)
