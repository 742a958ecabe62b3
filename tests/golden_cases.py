"""Decks and deterministic inputs shared by tests/golden/make_golden.py (which runs the compiled reference on them)
and by the tests (which run the oracle and the CUDA path on the same decks).  Test infrastructure only."""
import os
import zlib

import numpy as np

from mc_old_b200 import decks

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
XS_DIR = os.path.join(ROOT, "oracle", "_ref", "xs_library")


def function_decks():
    """Decks whose object graphs the function-level vectors are taken from."""
    return {
        "heu": decks.heu_sphere(samples=100, active=1, passive=1),
        "gcr": decks.gcr(samples=100, active=1, passive=1),
        "shield": decks.shielding(samples=100),
        "fsf": decks.fixed_source_fissile(samples=100),
        "ucube": decks.ucube(samples=100, active=1, passive=1),
    }


def run_decks():
    """name -> (xml, needs the patched reference build).  Sizes keep every reference run under a few seconds."""
    return {
        "heu": (decks.heu_sphere(samples=2000, active=3, passive=2), False),
        "heu_tallies": (decks.heu_sphere(samples=1500, active=3, passive=2, estimators=True), False),
        "heu_entropy": (decks.heu_sphere(samples=1000, active=2, passive=2, entropy=True), False),
        "ucube": (decks.ucube(samples=1000, active=3, passive=2), False),
        "gcr": (decks.gcr(samples=100, active=2, passive=2), False),
        "gcr_trmm": (decks.gcr(samples=60, active=2, passive=1, trmm=True), False),
        "slab": (decks.slab(samples=20000), True),
        "shield": (decks.shielding(samples=4000), True),
        "shield_split": (decks.shielding(samples=3000, split=True), True),
        "fsf": (decks.fixed_source_fissile(samples=2000), True),
        "fsf_comb": (decks.fixed_source_fissile(samples=2000, comb=(3, 2)), True),
        "leak_time": (decks.heu_leakage(samples=3000), False),
        "gcr_td": (decks.gcr_td(samples=300), True),
        "gcr_td_comb": (decks.gcr_td(samples=200, linear="1e-8 6 2e-5", comb=(8, 4), groups=10), True),
    }


# nuclides of each function deck's materials, for grid points (ZAID files)
_MAT_ZAIDS = {
    "heu": [["092235", "092238"]],
    "gcr": [["092235", "092238", "008016", "006000"]],
    "shield": [["005010", "005011", "006000"], ["001001", "008016"], ["002003"]],
    "fsf": [["092235", "006000"], ["006000"]],
    "ucube": [["092235"]],
}
_NUC_ZAIDS = {
    "heu": ["092235", "092238"], "gcr": ["092235", "092238", "008016", "006000"],
    "shield": ["008016", "001001", "002003", "006000", "005010", "005011"], "fsf": ["092235", "006000"], "ucube": ["092235"],
}


def _rng(*key):
    return np.random.default_rng(zlib.crc32(repr(key).encode()))


def grid_energies(zaid):
    """First column of xs_library/<ZAID>.txt after the 4 header lines."""
    out = []
    with open(os.path.join(XS_DIR, zaid + ".txt")) as f:
        for i, line in enumerate(f):
            if i < 4:
                continue
            t = line.split()
            if len(t) >= 5:
                out.append(float(t[0]))
    return np.array(out)


def energies(name, m):
    """Log-uniform energies over and beyond the tables' range + grid points of the material's nuclides and their
    +-1 ulp neighbours + the range ends.  m >= 100 addresses nuclide m-100 of the deck instead of a material."""
    r = _rng("E", name, m)
    zaids = _MAT_ZAIDS[name][m] if m < 100 else [_NUC_ZAIDS[name][m - 100]]
    parts = [np.exp(r.uniform(np.log(1e-6), np.log(3e8), 1500))]
    for z in zaids:
        g = grid_energies(z)
        pick = g[r.integers(0, g.size, 250)]
        parts += [pick, np.nextafter(pick, 0.0), np.nextafter(pick, np.inf), g[:2], g[-2:]]
        dup = g[1:][g[1:] == g[:-1]][:50]  # duplicate grid energies (zero-width bins)
        parts += [dup, np.nextafter(dup, 0.0), np.nextafter(dup, np.inf)]
    parts.append(np.array([1e-5, 2e7, 1.0, 1e6, 14e6, 0.0253]))
    return np.concatenate(parts)


def select_inputs(name, m, kind):
    r = _rng("S", name, m, kind)
    E = energies(name, m)[:1200]
    xi = r.uniform(0.0, 1.0, E.size)
    xi[:8] = [0.0, 1.0 - 2.0 ** -53, 0.5, 2.0 ** -63, 0.999999, 1e-12, 0.25, 0.75]
    return E, xi


def geometry_inputs(name):
    r = _rng("G", name)
    n = 200
    pos = r.uniform(-8.0, 10.0, (n, 3))
    d = r.normal(size=(n, 3))
    d /= np.linalg.norm(d, axis=1)[:, None]
    d[:6] = [[1, 0, 0], [0, 1, 0], [0, 0, 1], [-1, 0, 0], [0, -1, 0], [0, 0, -1]]  # axis-parallel flights
    pos[6] = [0.0, 0.0, 0.0]
    return pos, d


def sample_inputs(name):
    r = _rng("X", name)
    n = 64
    seeds = r.integers(1, 2 ** 62, n, dtype=np.uint64) | np.uint64(1)
    E = np.exp(r.uniform(np.log(1e-3), np.log(2e7), n))
    E[:4] = [0.0253, 1.0, 1e6, 14e6]
    d = r.normal(size=(n, 3))
    d /= np.linalg.norm(d, axis=1)[:, None]
    return seeds, E, d
