#!/usr/bin/env python3
"""Where the wall time of a tiny MCB.exe run goes outside main(): spawn -> main (dynamic linking, fatbin registration),
main's phases (MCB_TIMING), teardown, end of main -> process gone."""
import os, subprocess, sys, tempfile, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mc_old_b200 import decks
exe = os.path.join(ROOT, "mc_old_b200", "MCB.exe")
env = dict(os.environ, MCB_XS_LIBRARY=os.path.join(ROOT, "data", "xs_library"), MCB_TIMING="1")
for rep in range(4):
    d = tempfile.mkdtemp(prefix="mcb_t_")
    decks.write(d, decks.slab(samples=100000))
    t0 = time.time()
    r = subprocess.run([exe, d], env=env, capture_output=True, text=True)
    t1 = time.time()
    lines = [l for l in r.stderr.splitlines() if l.startswith("[mcb timing]")]
    w = lines[-1].split()
    e_main, e_end = float(w[4]), float(w[7])
    print("run %d: wall %.3f s | spawn->main %.3f | %s | %s | end of main -> exit %.3f" % (rep, t1 - t0, e_main - t0, lines[0][13:], lines[-1].split("(")[1].rstrip(")"), t1 - e_end), flush=True)
