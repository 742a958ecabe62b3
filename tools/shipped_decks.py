#!/usr/bin/env python3
"""The reference's example decks at the sizes they ship with, as a user runs them: `MCB.exe <dir>` (deck load, xs_library
parse, every cycle, output.h5) in wall seconds, next to the compiled reference `oracle/_ref/MC_ref <dir>` on the same
directory on this machine's host cores (one thread: the reference is single-threaded).  The decks are written by
mc_old_b200/decks.py (the examples' physics, history and cycle counts as shipped; /root/reference is not on the GPU box).

    tools/shipped_decks.py [--ref-timeout S] [--skip-ref]      -> one line per deck + a JSON summary on the last line
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mc_old_b200 import decks  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--ref-timeout", type=float, default=90.0)
ap.add_argument("--skip-ref", action="store_true")
ap.add_argument("--only", default="")
args = ap.parse_args()

# name: (deck text, histories in the whole run, reference: True = MC_ref, "patched" = MC_ref_patched (the unmodified
# reference dereferences a null pointer on capture-only nuclides, SURVEY patch A), False = its loader rejects the deck)
CASES = {
    "HEU_sphere_criticality (1e4 x 200, entropy)": (decks.heu_sphere(samples=10000, active=180, passive=20, entropy=True), 10000 * 200, True),
    "UCube (1e5 x 115)": (decks.ucube(samples=100000, active=100, passive=15), 100000 * 115, True),
    "slab_analytic (1e7)": (decks.slab(samples=10_000_000), 10_000_000, "patched"),
    "infinite_GCR_TRMM (400 x 110, 3500 tallies)": (decks.gcr(samples=400, active=100, passive=10, trmm=True), 400 * 110, True),
    "HEU_sphere_leakage (1e4)": (decks.heu_leakage(samples=10000), 10000, True),
    "infinite_GCR_TD (2500)": (decks.gcr_td(samples=2500), 2500, True),
    "sphere_detection (1e7, <disk_z>)": (decks.sphere_detection(samples=10_000_000), 10_000_000, False),
    "shielding_vReduction (1e8)": (decks.shielding(samples=100_000_000), 100_000_000, "patched"),
}
exe = os.path.join(ROOT, "mc_old_b200", "MCB.exe")
ref_dir = os.path.join(ROOT, "oracle", "_ref")
ref = os.path.join(ref_dir, "MC_ref")
env = dict(os.environ, MCB_XS_LIBRARY=os.path.join(ROOT, "data", "xs_library"), MCB_TIMING="1")
out = []
# one untimed run first: CUDA context creation and module load of a fresh process are paid by every run below too,
# but the very first touch of the driver on a fresh box is not the program's
d0 = tempfile.mkdtemp(prefix="mcb_ship_")
decks.write(d0, decks.slab(samples=1000))
subprocess.run([exe, d0], env=env, capture_output=True)
for name, (xml, n_hist, ref_ok) in CASES.items():
    if args.only and args.only not in name:
        continue
    d = tempfile.mkdtemp(prefix="mcb_ship_")
    decks.write(d, xml)
    t = time.perf_counter()
    r = subprocess.run([exe, d], env=env, capture_output=True, text=True)
    t_gpu = time.perf_counter() - t
    ok = r.returncode == 0 and os.path.exists(os.path.join(d, "output.h5"))
    rec = {"deck": name, "histories": n_hist, "mcb_exe_wall_s": round(t_gpu, 3), "mcb_exe_ok": ok}
    for line in r.stderr.splitlines():   # [mcb timing] deck .. s  context .. s  first cycle .. s  other N cycles .. s  output .. s
        if line.startswith("[mcb timing]"):
            w = line.split()
            rec["mcb_exe_phases_s"] = {"deck_and_xs_library": float(w[3]), "cuda_start_and_context": float(w[6]),
                                       "all_cycles": round(float(w[10]) + float(w[15]), 4), "output_h5": float(w[18])}
    if not ok:
        rec["mcb_exe_tail"] = (r.stdout + r.stderr)[-300:]
    if not args.skip_ref and os.path.exists(ref):
        if not ref_ok:
            rec["mc_ref"] = "the reference's loader rejects this deck (setup.cpp:1051-1063)"
        else:
            d2 = tempfile.mkdtemp(prefix="mcb_ship_ref_")
            decks.write(d2, xml)
            if ref_ok == "patched":
                rec["mc_ref_build"] = "MC_ref_patched (capture-only nuclides crash the unmodified reference, SURVEY patch A)"
            # 1e8 histories take the reference several minutes: it runs 1/100 of them and the time is scaled (a fixed-source
            # run is linear in its histories), which the record says
            scale = 100 if n_hist >= 100_000_000 else 1
            if scale > 1:
                decks.write(d2, xml.replace('samples="%g"' % n_hist, 'samples="%g"' % (n_hist // scale)))
                assert 'samples="%g"' % (n_hist // scale) in open(os.path.join(d2, "input.xml")).read()
            t = time.perf_counter()
            try:
                rr = subprocess.run([ref + "_patched" if ref_ok == "patched" else ref, d2], cwd=ref_dir, capture_output=True, text=True, timeout=args.ref_timeout)
                t_ref = time.perf_counter() - t
                rec["mc_ref_ok"] = rr.returncode == 0
                if scale > 1:
                    rec["mc_ref_measured"] = "%.3f s for %g histories, scaled x%d" % (t_ref, n_hist // scale, scale)
                rec["mc_ref_wall_s"] = round(t_ref * scale, 3)
                if rr.returncode == 0:
                    rec["speedup_wall"] = round(rec["mc_ref_wall_s"] / t_gpu, 1)
            except subprocess.TimeoutExpired:
                rec["mc_ref"] = "still running after %.0f s (stopped)" % args.ref_timeout
                rec["speedup_wall_at_least"] = round(args.ref_timeout / t_gpu, 1)
    print(json.dumps(rec), flush=True)
    out.append(rec)
print(json.dumps({"shipped_decks": out, "host_cores_used_by_reference": 1}))
