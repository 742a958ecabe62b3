"""Time-dependent mode (<tdmc>; setup.cpp:133-169, general.cpp:187-195, time_dependent.cpp, fixed_source.cpp:25-40,
FilterTDMC Estimator.cpp:247-263).  The oracle's bit-parity with the compiled reference on the two TDMC decks is in
test_oracle_golden.py::test_whole_run_bit_exact[gcr_td*], the GPU's parity with the oracle in
test_gpu_transport.py::test_tallies_history_parity[gcr_td*]; here: the deck side and the host program."""
import os
import subprocess

import numpy as np
import pytest

import h5mini
import mc_old_b200 as mcb
from mc_old_b200 import decks
import oracle_lib as ol

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_census_grid_as_the_reference_builds_it():
    """time="...": interval j = time j minus the previous INTERVAL (setup.cpp:146), which is the previous time only for the
    first two; time_linear="a n b": a, then steps of (b - a) / n by repeated addition, the last point replaced by b"""
    t, dt = mcb.Deck(xml=decks.gcr_td(samples=10)).tdmc()
    assert np.array_equal(t, [3e-8, 15e-8, 4e-6, 100e-6])
    want = [3e-8 - 0.0]
    for s in t[1:]:
        want.append(s - want[-1])
    assert np.array_equal(dt, want) and dt[2] != t[2] - t[1]
    t, dt = mcb.Deck(xml=decks.gcr_td(samples=10, linear="1e-8 6 2e-5")).tdmc()
    step = (2e-5 - 1e-8) / 6
    grid = [1e-8]
    while grid[-1] < 2e-5:
        grid.append(grid[-1] + step)
    grid[-1] = 2e-5
    assert np.array_equal(t, grid) and len(dt) == len(t)
    assert mcb.Deck(xml=decks.slab(samples=10)).tdmc()[0].size == 0


def test_deck_of_a_time_dependent_run():
    deck = mcb.Deck(xml=decks.gcr_td(samples=10))
    assert deck.mode == "time-dependent" and deck.info["ksearch"] == 0 and deck.info["n_cycle"] == 1
    spectrum, rates = deck.estimators()
    # the <tdmc/> filter comes first in the index order (setup.cpp:703-708) and is reported as "time"
    assert [f["size"] for f in spectrum["filters"]] == [4, 1, 50] and spectrum["n_tallies"] == 200
    assert [f["size"] for f in rates["filters"]] == [1]
    with pytest.raises(ValueError, match="ksearch and tdmc could not coexist"):
        mcb.Deck(xml=decks.gcr(samples=10).replace("</simulation>", '<tdmc time="1.0 2.0"/></simulation>'))


def test_oracle_history_mode_equals_global_mode_statistically():
    """the per-history stream layout the GPU uses against the reference's single stream: same physics, different draws"""
    deck = mcb.Deck(xml=decks.gcr_td(samples=1500, groups=5))
    out = []
    for mode, pick in ((ol.RNG_GLOBAL, ol.PICK_CDF), (ol.RNG_HISTORY, ol.PICK_FLOOR)):
        o = ol.Oracle(deck, rng_mode=mode, pick_mode=pick)
        o.run()
        out.append(o.tallies())
    (m1, u1), (m2, u2) = out
    s = np.sqrt(u1 * u1 + u2 * u2)
    ok = s > 0
    assert ok.sum() >= 8
    assert np.all(np.abs(m1 - m2)[ok] <= 4.5 * s[ok])
    assert np.mean(np.abs(m1 - m2)[ok] <= 2.0 * s[ok]) >= 0.8


@pytest.mark.gpu
def test_host_program_runs_a_time_dependent_deck(tmp_path):
    d = str(tmp_path)
    decks.write(d, decks.gcr_td(samples=20000))
    env = dict(os.environ, MCB_XS_LIBRARY=mcb.default_xs_dir())
    out = subprocess.run([os.path.join(ROOT, "mc_old_b200", "MCB.exe"), d], capture_output=True, text=True, env=env, timeout=600)
    assert out.returncode == 0, out.stdout + out.stderr
    f = h5mini.File(os.path.join(d, "output.h5"))
    assert f.root["summary/mode"].value == "time-dependent"
    assert np.array_equal(f.root["summary/tdmc/time"].value, [3e-8, 15e-8, 4e-6, 100e-6])
    assert f.root["spectrum"].attrs["indexing"] == "[time][cell][energy]"
    flux = f.root["spectrum/flux/mean"].value
    assert flux.shape == (4, 1, 50)
    # the census spectra match the oracle run of the same deck (reference stream layout) within the statistics of both
    deck = mcb.Deck(xml=decks.gcr_td(samples=2000))
    o = ol.Oracle(deck, rng_mode=ol.RNG_GLOBAL, pick_mode=ol.PICK_CDF)
    o.run()
    om, ou = o.tallies()
    gm, gu = flux.ravel(), f.root["spectrum/flux/uncertainty"].value.ravel()
    s = np.sqrt(ou[:200] ** 2 + gu ** 2)
    ok = (ou[:200] > 0) & (gu > 0)
    assert ok.sum() > 40
    z = np.abs(gm - om[:200])[ok] / s[ok]
    assert np.all(z < 5.0) and np.mean(z < 2.0) > 0.85
