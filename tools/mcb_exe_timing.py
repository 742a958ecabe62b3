#!/usr/bin/env python3
"""MCB.exe phase timing (MCB_TIMING=1) on a few shipped-size decks: where the wall time of a small run goes."""
import os, subprocess, sys, tempfile, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mc_old_b200 import decks
exe = os.path.join(ROOT, "mc_old_b200", "MCB.exe")
env = dict(os.environ, MCB_XS_LIBRARY=os.path.join(ROOT, "data", "xs_library"), MCB_TIMING="1", MCB_TRACE_HOST="1")
cases = {"warm-up slab 1e3": decks.slab(samples=1000), "slab 1e7": decks.slab(samples=10_000_000), "sphere_detection 1e7": decks.sphere_detection(samples=10_000_000),
         "gcr_td 2500": decks.gcr_td(samples=2500), "gcr_trmm 400 x 110": decks.gcr(samples=400, active=100, passive=10, trmm=True),
         "leakage 1e4": decks.heu_leakage(samples=10000), "heu 1e4 x 200": decks.heu_sphere(samples=10000, active=180, passive=20, entropy=True),
         "shielding 1e8": decks.shielding(samples=100_000_000)}
for name, xml in cases.items():
    d = tempfile.mkdtemp(prefix="mcb_t_")
    decks.write(d, xml)
    t = time.perf_counter()
    r = subprocess.run([exe, d] + sys.argv[1:], env=env, capture_output=True, text=True)
    print("%-22s wall %.3f s  rc %d  %s" % (name, time.perf_counter() - t, r.returncode, "\n   " + "\n   ".join(l for l in r.stderr.strip().splitlines() if "create" in l or "timing" in l)), flush=True)
