import sys, time
sys.path.insert(0, '/root/repo')
import mc_old_b200 as mcb
from mc_old_b200 import decks
for n in (1_000_000, 4_000_000):
    deck = mcb.Deck(xml=decks.heu_leakage(samples=n)); deck.set_run(n_cycle=2, n_passive=0)
    ctx = mcb.Context(deck)
    for c in range(2):
        t = time.perf_counter(); r = ctx.run_cycle(); dt = time.perf_counter() - t
    print(n, "%.1f ms  %.3e hist/s %.3e tracks/s" % (dt * 1e3, n / dt, r.n_tracks / dt), flush=True)
    ctx.close()
