#!/usr/bin/env python3
"""Generates the committed golden fixtures from the COMPILED REFERENCE (oracle/_ref, built from /root/reference by
oracle/build_ref.py).  Run in the container where /root/reference exists:

    python tests/golden/make_golden.py

Outputs (small, committed):
  tests/golden/functions.npz   function-level vectors: macroscopic cross sections, per-nuclide micro data, nuclide
                               selection with injected xi, surface eval/distance/reflect, free-gas scatter and Watt
                               samples from given LCG seeds — all computed by the reference's own objects
                               (libref_harness.so)
  tests/golden/runs.json       whole-run outputs of MC_ref / MC_ref_patched (k per cycle, entropy, Ntrack, every
                               tally mean/uncertainty), floats stored as C99 hex strings so they round-trip exactly

Nothing here is used by the product; tests/ compare the oracle (oracle/mc_oracle.c) and the CUDA path with it.
"""
import json
import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import oracle_lib as ol  # noqa: E402
from mc_old_b200 import decks  # noqa: E402
import golden_cases as gc  # noqa: E402


def runs_only(names):
    """regenerate only the named whole-run records of runs.json (python make_golden.py --runs name ...)"""
    path = os.path.join(HERE, "runs.json")
    runs = json.load(open(path))
    for name in names:
        xml, patched = gc.run_decks()[name]
        d = decks.write(tempfile.mkdtemp(prefix="gold_"), xml)
        stdout, parsed = ol.run_ref(d, patched=patched)
        rec = {"patched": patched, "stdout_cycle_lines": [ln for ln in stdout.splitlines() if ln[:1].isdigit()]}
        for k, v in parsed.items():
            rec[k] = v if isinstance(v, str) else ([int(x) for x in v] if v.dtype == np.uint64 else [float(x).hex() for x in v])
        runs[name] = rec
    with open(path, "w") as f:
        json.dump(runs, f, indent=0, sort_keys=True)
    print("updated", names)


def main():
    assert ol.have_ref(), "oracle/_ref is not built"
    if len(sys.argv) > 2 and sys.argv[1] == "--runs":
        return runs_only(sys.argv[2:])
    fn = {}
    for name, xml in gc.function_decks().items():
        d = decks.write(tempfile.mkdtemp(prefix="gold_"), xml)
        sim = ol.RefSim(d)
        counts = np.zeros(5, dtype=np.int32)
        sim.L.refh_counts(sim.h, counts.ctypes.data)
        n_nuc, n_mat, n_surf = int(counts[0]), int(counts[1]), int(counts[2])
        for m in range(n_mat):
            E = gc.energies(name, m)
            fn[f"xs/{name}/{m}/E"] = E
            fn[f"xs/{name}/{m}/sigma"] = sim.sigma(m, E)
            for kind in (0, 1):
                Es, xi = gc.select_inputs(name, m, kind)
                fn[f"select/{name}/{m}/{kind}/E"] = Es
                fn[f"select/{name}/{m}/{kind}/xi"] = xi
                fn[f"select/{name}/{m}/{kind}/out"] = sim.select(m, kind, Es, xi)
        for n in range(n_nuc):
            E = gc.energies(name, 100 + n)[:600]
            fn[f"micro/{name}/{n}/E"] = E
            fn[f"micro/{name}/{n}/out"] = sim.micro(n, E)
        pos, dirs = gc.geometry_inputs(name)
        out = np.zeros((n_surf, pos.shape[0], 6))
        for s in range(n_surf):
            for i in range(pos.shape[0]):
                out[s, i] = sim.surface(s, pos[i], dirs[i])
        fn[f"surface/{name}/pos"] = pos
        fn[f"surface/{name}/dir"] = dirs
        fn[f"surface/{name}/out"] = out
        # sampling from explicit LCG seeds (libm-dependent: bit-exact for the oracle on x86-64, tolerance on the GPU)
        seeds, Ein, dirs_in = gc.sample_inputs(name)
        sc = np.zeros((n_nuc, seeds.size, 5)); sc_seed = np.zeros((n_nuc, seeds.size), dtype=np.uint64)
        wt = np.zeros((n_nuc, seeds.size)); wt_seed = np.zeros((n_nuc, seeds.size), dtype=np.uint64)
        for n in range(n_nuc):
            for i, s in enumerate(seeds):
                io, s1 = sim.scatter_sample(n, dirs_in[i], Ein[i], int(s))
                sc[n, i] = io; sc_seed[n, i] = s1
                if gc._NUC_ZAIDS[name][n].startswith("092"):  # Watt parameters exist for fissile nuclides only
                    v, s2 = sim.watt(n, Ein[i], int(s))
                    wt[n, i] = v; wt_seed[n, i] = s2
        fn[f"sample/{name}/seeds"] = seeds; fn[f"sample/{name}/E"] = Ein; fn[f"sample/{name}/dir"] = dirs_in
        fn[f"scatter/{name}/out"] = sc; fn[f"scatter/{name}/seed"] = sc_seed
        fn[f"watt/{name}/out"] = wt; fn[f"watt/{name}/seed"] = wt_seed
    np.savez_compressed(os.path.join(HERE, "functions.npz"), **fn)

    runs = {}
    for name, (xml, patched) in gc.run_decks().items():
        d = decks.write(tempfile.mkdtemp(prefix="gold_"), xml)
        stdout, parsed = ol.run_ref(d, patched=patched)
        rec = {"patched": patched, "stdout_cycle_lines": [ln for ln in stdout.splitlines() if ln[:1].isdigit()]}
        for k, v in parsed.items():
            if isinstance(v, str):
                rec[k] = v
            elif v.dtype == np.uint64:
                rec[k] = [int(x) for x in v]
            else:
                rec[k] = [float(x).hex() for x in v]
        runs[name] = rec
        print(name, "ok", {k: len(v) for k, v in rec.items() if isinstance(v, list)} if False else "")
    with open(os.path.join(HERE, "runs.json"), "w") as f:
        json.dump(runs, f, indent=0, sort_keys=True)
    print("wrote functions.npz (%d arrays) and runs.json (%d runs)" % (len(fn), len(runs)))


if __name__ == "__main__":
    main()
