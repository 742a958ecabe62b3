"""The oracle (oracle/mc_oracle.c) against the reference: committed golden vectors produced by the compiled
reference (tests/golden/make_golden.py), the known-answer vectors of the reference's own tests, and — where
oracle/_ref is present — the live reference binary.  CPU only.

Bar: bit-exact.  The oracle is x86-64 double arithmetic without FMA contraction, like the reference build, and uses
the same libm, so even the sampled kinematics (free-gas scatter, Watt) and whole runs agree bit for bit.
"""
import os
import tempfile

import numpy as np
import pytest

import golden_cases as gc
import mc_old_b200 as mcb
import oracle_lib as ol
from mc_old_b200 import decks

FN_DECKS = gc.function_decks()
RUN_DECKS = gc.run_decks()


def _mats(golden_fn, name):
    return sorted({int(k.split("/")[2]) for k in golden_fn.files if k.startswith("xs/%s/" % name)})


def _nucs(golden_fn, name):
    return sorted({int(k.split("/")[2]) for k in golden_fn.files if k.startswith("micro/%s/" % name)})


# ---- known answers from the reference's own tests -------------------------------------------------
def test_binary_search_known_answers():
    """test/test_Solver.cpp:78-109"""
    L = ol.oracle()
    v = np.array([2.0, 3.0, 4.0, 5.0, 6.0, 7.0, 8.0, 9.0, 10.0])
    for x, want in [(4.3, 2), (7.3, 5), (4.0, 1), (7.0, 4), (2.0, -1), (10.0, 7), (1.0, -1), (12.0, 8)]:
        assert L.mco_binary_search(x, v.ctypes.data, v.size) == want
    w = np.array([0.0, 5.0, 5.0, 10.0])
    assert L.mco_binary_search(1.0, w.ctypes.data, 4) == 0
    assert L.mco_binary_search(8.0, w.ctypes.data, 4) == 2


def test_interpolate_known_answers():
    """test/test_Solver.cpp:9-31: the three calls all give 3.0"""
    L = ol.oracle()
    assert L.mco_interpolate(3.0, 1.0, 5.0, 1.0, 5.0) == 3.0
    assert L.mco_interpolate(2.0, 1.0, 3.0, 2.0, 4.0) == 3.0
    assert L.mco_interpolate(1.5, 1.0, 2.0, 2.0, 4.0) == 3.0


def test_geometry_quad_known_answers():
    """test/test_Solver.cpp:33-76"""
    L = ol.oracle()
    MAXF = 3.402823466e+38
    for a, b, c in [(10.0, 1.0, 10.0), (1.0, 2.0, 1.0), (1.0, 3.0, 2.0)]:
        assert L.mco_geometry_quad(a, b, c) >= 0.99 * MAXF
    assert L.mco_geometry_quad(1.0, -3.0, 2.0) == 1.0
    assert L.mco_geometry_quad(1.0, -1.0, -2.0) == 2.0


def test_lcg_known_answers():
    """RN_CHECK of src/Random.cpp:111-117: states #1-5 and #123456-123460 of the generator from seed 1"""
    L = ol.oracle()
    check = [3512401965023503517, 5461769869401032777, 1468184805722937541, 5160872062372652241, 6637647758174943277,
             794206257475890433, 4662153896835267997, 6075201270501039433, 889694366662031813, 7299299962545529297]
    s = 1
    got = []
    for i in range(1, 123461):
        s = L.mco_lcg_next(s)
        if i <= 5 or i >= 123456:
            got.append(s)
    assert got == check
    # skip-ahead (Random.cpp:130-149) lands on the same states
    assert [L.mco_lcg_skip(1, n) for n in (1, 2, 3, 4, 5, 123456, 123460)] == check[:5] + [check[5], check[9]]
    assert L.mco_lcg_skip(1, 0) == 1


def test_scatter_direction_known_answers():
    """test/test_Solver.cpp:111-166 with the first four Urand() of seed 1 (tolerance 1e-6, as the reference's Approx)"""
    L = ol.oracle()
    s = 1
    xi = []
    for _ in range(4):
        s = L.mco_lcg_next(s)
        xi.append(s * 2.0 ** -63)
    assert np.allclose(xi, [0.380815, 0.592166, 0.159181, 0.559543], atol=1e-6)
    cases = [((2 / 3, 2 / 3, 1 / 3), 0.4, (-0.332776, 0.549648, 0.766255)),
             ((2 / 3, 2 / 3, 1 / 3), -0.4, (-0.092800, -0.802140, 0.589881)),
             ((0, 0, 1), 0.4, (-0.771302, -0.495069, 0.4)),
             ((0, 0, 1), -0.4, (0.334943, 0.853120, -0.4))]
    for (d, mu, want), x in zip(cases, xi):
        din = np.array(d, dtype=np.float64)
        out = np.zeros(3)
        L.mco_scatter_direction(din.ctypes.data, mu, x, out.ctypes.data)
        assert np.allclose(out, want, atol=2e-6), (d, mu, out, want)


def test_particle_speed_constants():
    """test/unit_test.cpp:37-90: speed = sqrt(E*191312955.067)*100, E(2200 m/s ...) = v^2 * 5.2270376e-13"""
    L = ol.oracle()
    assert L.mco_speed_of_energy(3.0) == pytest.approx(np.sqrt(3.0 * 191312955.067) * 100.0, rel=1e-9)
    assert L.mco_energy_of_speed(2200.0) == 2200.0 * 2200.0 * 5.2270376e-13 or \
        L.mco_energy_of_speed(2200.0) == pytest.approx(2200.0 ** 2 * 5.2270376e-13, rel=1e-15)


# ---- golden vectors of the compiled reference ------------------------------------------------------------------
@pytest.mark.parametrize("name", sorted(FN_DECKS))
def test_xs_lookup_bit_exact(name, golden_fn, deck_cache):
    """Material::Sigma{T,S,C,F}, nuSigmaF (Material.cpp:18-65) incl. grid points +-1 ulp, duplicates, clamps"""
    deck = deck_cache(FN_DECKS[name])
    for m in _mats(golden_fn, name):
        E = golden_fn["xs/%s/%d/E" % (name, m)]
        want = golden_fn["xs/%s/%d/sigma" % (name, m)]
        got = ol.xs_lookup(deck, m, E)
        assert np.array_equal(got, want[:, :5]), "material %d of %s" % (m, name)


@pytest.mark.parametrize("name", sorted(FN_DECKS))
def test_select_channel_bit_exact(name, golden_fn, deck_cache):
    """Material::nuclide_scatter / nuclide_nufission (Material.cpp:106-125) with injected xi"""
    deck = deck_cache(FN_DECKS[name])
    for m in _mats(golden_fn, name):
        for kind in (0, 1):
            key = "select/%s/%d/%d" % (name, m, kind)
            got = ol.select_channel(deck, m, kind, golden_fn[key + "/E"], golden_fn[key + "/xi"])
            assert np.array_equal(got, golden_fn[key + "/out"]), key


@pytest.mark.parametrize("name", sorted(FN_DECKS))
def test_beta_bit_exact(name, golden_fn, deck_cache):
    """Nuclide::beta (Nuclide.cpp:74-77)"""
    deck = deck_cache(FN_DECKS[name])
    for n in _nucs(golden_fn, name):
        E = golden_fn["micro/%s/%d/E" % (name, n)]
        want = golden_fn["micro/%s/%d/out" % (name, n)][:, 5]
        assert np.array_equal(ol.beta(deck, n, E), want)


@pytest.mark.parametrize("name", sorted(FN_DECKS))
def test_surfaces_bit_exact(name, golden_fn, deck_cache):
    """Surface::eval / distance / reflect (Geometry.cpp:29-222)"""
    deck = deck_cache(FN_DECKS[name])
    L = ol.oracle()
    pos, dirs, want = (golden_fn["surface/%s/%s" % (name, k)] for k in ("pos", "dir", "out"))
    for s in range(want.shape[0]):
        for i in range(pos.shape[0]):
            p = np.ascontiguousarray(pos[i]); d = np.ascontiguousarray(dirs[i])
            assert L.mco_surface_eval(deck.problem, s, p.ctypes.data) == want[s, i, 0]
            assert L.mco_surface_distance(deck.problem, s, p.ctypes.data, d.ctypes.data) == want[s, i, 1]
            r = d.copy()
            L.mco_surface_reflect(deck.problem, s, r.ctypes.data)
            assert np.array_equal(r, want[s, i, 2:5])


@pytest.mark.parametrize("name", sorted(FN_DECKS))
def test_scatter_and_watt_samples_bit_exact(name, golden_fn, deck_cache):
    """ReactionScatter::sample (Reaction.cpp:27-118) and DistributionWatt::sample (Distribution.cpp:34-73) from given
    LCG seeds: same outputs and same number of draws (final seed)"""
    import ctypes as C
    deck = deck_cache(FN_DECKS[name])
    L = ol.oracle()
    seeds, E, dirs = (golden_fn["sample/%s/%s" % (name, k)] for k in ("seeds", "E", "dir"))
    sc, sc_seed = golden_fn["scatter/%s/out" % name], golden_fn["scatter/%s/seed" % name]
    wt, wt_seed = golden_fn["watt/%s/out" % name], golden_fn["watt/%s/seed" % name]
    for n in range(sc.shape[0]):
        for i in range(seeds.size):
            s = C.c_uint64(int(seeds[i]))
            io = np.array([dirs[i, 0], dirs[i, 1], dirs[i, 2], E[i], 0.0])
            L.mco_scatter_sample(deck.problem, n, C.byref(s), io.ctypes.data)
            assert np.array_equal(io, sc[n, i]) and s.value == int(sc_seed[n, i]), (name, n, i)
            if wt_seed[n, i]:
                s = C.c_uint64(int(seeds[i]))
                v = L.mco_watt_sample(deck.problem, n, C.byref(s), float(E[i]))
                assert v == wt[n, i] and s.value == int(wt_seed[n, i])


def _tally_arrays(rec):
    """tally means / uncertainties of a golden run in estimator order -> flat vectors in the C-ABI's order"""
    means = [(k, v) for k, v in rec.items() if k.endswith("/mean") and not k.startswith("/ksearch")]
    return means


@pytest.mark.parametrize("name", sorted(RUN_DECKS))
def test_whole_run_bit_exact(name, golden_runs, deck_cache):
    """Simulator::start() (handler.cpp:11-48) on every deck: k per cycle, running k, entropy, Ntrack and every tally
    mean / uncertainty equal the compiled reference's output bit for bit (global RNG stream, CDF source pick)."""
    import mc_old_b200 as mcb
    xml, _ = RUN_DECKS[name]
    rec = golden_runs[name]
    deck = mcb.Deck(xml=xml)
    o = ol.Oracle(deck, rng_mode=ol.RNG_GLOBAL, pick_mode=ol.PICK_CDF)
    res = o.run()
    assert sum(r.n_tracks for r in res) == int(rec["/summary/Ntrack"][0])
    if deck.info["ksearch"]:
        assert np.array_equal(np.array([r.k_cycle for r in res]), rec["/ksearch/k_cycle"])
        assert np.array_equal(np.array([r.H for r in res]), rec["/ksearch/H_cycle"])
        act = [r for r in res[deck.info["n_passive"]:]]
        assert np.array_equal(np.array([r.k_avg for r in act]), rec["/ksearch/k_active/mean"])
        assert np.array_equal(np.array([r.k_uncer for r in act]), rec["/ksearch/k_active/uncertainty"])
    mean, uncer = o.tallies()
    # the reference writes /<estimator>/<score>/{mean,uncertainty}; our flat order is [estimator][score][bins]
    import report_order
    want_mean, want_uncer = report_order.flatten(deck, rec)
    keep = np.ones(mean.size, dtype=bool)
    for est in deck.estimators():
        # TRM_simple's NuFissionDelayed* scores read the REAL particle's energy_old, which the reference leaves
        # uninitialised at birth (Particle ctor calls set_energy before p_energy is set, include/Particle.h:30-34):
        # its values there are stack garbage.  The oracle defines E_old = E at birth; those bins are not comparable.
        bins = est["n_tallies"] // max(len(est["scores"]), 1)
        for k, score in enumerate(est["scores"]):
            if est["name"] == "TRM_simple" and score.startswith("NuFissionDelayed"):
                keep[est["tally_begin"] + k * bins: est["tally_begin"] + (k + 1) * bins] = False
    assert keep.sum() >= mean.size - 240
    assert np.array_equal(mean[keep], want_mean[keep])
    assert np.array_equal(uncer[keep], want_uncer[keep])


@pytest.mark.skipif(not ol.have_ref(), reason="oracle/_ref not built")
def test_live_reference_matches_golden(golden_runs):
    """the committed fixture is what the reference binary prints today (guards against a stale fixture)"""
    xml, patched = RUN_DECKS["heu"]
    d = decks.write(tempfile.mkdtemp(prefix="mcb_t_"), xml)
    _, parsed = ol.run_ref(d, patched=patched)
    assert np.array_equal(parsed["/ksearch/k_cycle"], golden_runs["heu"]["/ksearch/k_cycle"])


def test_slab_analytic(golden_runs):
    """test/test_integral_Simulator.cpp:20-29: leakage = exp(-(1.2*1 + 0.75*4)) = 0.0149956; the reference asks for
    1 sigma at its seed, we ask for 3 sigma of the 20000-history golden run"""
    rec = golden_runs["slab"]
    m = rec["/leak_rate/cross/mean"][0]; u = rec["/leak_rate/cross/uncertainty"][0]
    assert abs(m - np.exp(-4.2)) < 3 * u


def test_sphere_detection_mcnp6_golden():
    """test/test_integral_Simulator.cpp:10-19: absorption in the He-3 tube of examples/sphere_detection = 6.9276e-5 per source
    particle (MCNP6), asked for within uncer + 1.15e-6.  The reference's loader rejects the deck's <disk_z> source
    (setup.cpp:1051-1063), so there is no reference run to pin the disk sampling to ("parity unpinned" for this one
    element); the MCNP6 number is the external anchor: 3 sigma of a 1e6-history oracle run + MCNP6's own 1.15e-6"""
    deck = mcb.Deck(xml=decks.sphere_detection(samples=1_000_000))
    assert deck.info["n_sources"] == 1 and deck.info["n_estimators"] == 1   # the second <estimators> block is never read
    orc = ol.Oracle(deck, rng_mode=ol.RNG_GLOBAL, pick_mode=ol.PICK_CDF)
    orc.run_cycle(); orc.end_simulation()
    m, u = orc.tallies()
    assert abs(m[1] - 6.9276e-5) <= 3 * u[1] + 1.15e-6, (m, u)
    assert 2e-6 < u[1] < 6e-6


def test_disk_source_needs_a_radius():
    with pytest.raises(Exception, match="radius"):
        mcb.Deck(xml=decks.sphere_detection(samples=10).replace('r="2.0" direction', 'r="0.0" direction'))


@pytest.mark.skipif(not ol.have_ref(), reason="oracle/_ref not built")
def test_search_cell_takes_the_first_match_in_deck_order_like_the_reference():
    """general.cpp:26-34 on overlapping cells: a void cell that overlaps the one-surface outside cell and precedes it in
    the deck catches every particle leaving the slab; they cross a far plane before they die.  The compiled reference
    (patched build: capture-only nuclides) and the oracle give the same two tallies bit for bit - the semantics the
    GPU's crossing shortcut has to preserve (tests/test_gpu_transport.py: slab_overlap)"""
    xml = decks.slab_overlap(samples=20000)
    d = decks.write(tempfile.mkdtemp(prefix="mcb_t_"), xml)
    _, parsed = ol.run_ref(d, patched=True)
    deck = mcb.Deck(xml=xml)
    orc = ol.Oracle(deck, rng_mode=ol.RNG_GLOBAL, pick_mode=ol.PICK_CDF)
    orc.run_cycle(); orc.end_simulation()
    m, u = orc.tallies()
    assert m[0] > 0 and m[1] == m[0]                                   # every leaked particle reaches the far plane
    assert parsed["/leak_rate/cross/mean"][0] == m[0] and parsed["/far_plane/cross/mean"][0] == m[1]
    assert parsed["/far_plane/cross/uncertainty"][0] == u[1]
