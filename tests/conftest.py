"""pytest configuration.

-m "not gpu": oracle vs the committed golden vectors of the reference, host-side logic (deck loader, sharding,
              close-out arithmetic, world-size-2 gloo runs), and that the C-ABI libraries load and export every symbol
              include/*.h declares.  No compute call touches a GPU.
-m gpu:       the parity tests proper — the CUDA path, called through the C-ABI, against the oracle / the golden
              vectors / size-independent properties.
"""
import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu")
    # build what is missing (the driver runs __graft_entry__.build() first; this is for a bare checkout)
    need = [os.path.join(ROOT, "mc_old_b200", "libmcbhost.so"), os.path.join(ROOT, "mc_old_b200", "libmcb200.so"),
            os.path.join(ROOT, "oracle", "libmc_oracle.so")]
    if not all(os.path.exists(p) for p in need):
        import __graft_entry__
        __graft_entry__.build()


def _has_gpu():
    try:
        import mc_old_b200 as mcb
        return mcb.cuda_lib().mcb_device_count() > 0
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _has_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def golden_fn():
    return np.load(os.path.join(ROOT, "tests", "golden", "functions.npz"))


@pytest.fixture(scope="session")
def golden_runs():
    with open(os.path.join(ROOT, "tests", "golden", "runs.json")) as f:
        raw = json.load(f)
    out = {}
    for name, rec in raw.items():
        d = {}
        for k, v in rec.items():
            if isinstance(v, list) and v and isinstance(v[0], str) and k != "stdout_cycle_lines":
                d[k] = np.array([float.fromhex(x) for x in v])
            elif isinstance(v, list) and k != "stdout_cycle_lines":
                d[k] = np.array(v, dtype=np.uint64)
            else:
                d[k] = v
        out[name] = d
    return out


@pytest.fixture(scope="session")
def deck_cache():
    """Decks are expensive to load (xs_library parsing): one instance per XML text."""
    import mc_old_b200 as mcb
    cache = {}

    def get(xml):
        if xml not in cache:
            cache[xml] = mcb.Deck(xml=xml)
        return cache[xml]
    return get
