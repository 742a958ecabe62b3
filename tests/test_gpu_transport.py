"""Run-level parity of the CUDA transport loop (mcb_run_cycle through the C-ABI) against the oracle.

Two kinds of comparison:

* TRAJECTORY parity against the oracle in MCO_RNG_HISTORY / MCO_PICK_FLOOR mode, which restates the GPU's stream
  layout (per-history streams RN_init_particle-style, per-site child streams) on top of the reference's physics:
  every history draws the same xi on both sides, so per-history k scores, the fission bank and the tallies agree
  to rounding of CUDA's libm vs glibc (log, sin, cos) — tolerance 1e-9 relative for >= 99.5 % of the histories
  (an ulp can flip a rejection test or a floor(), which changes that one history).
* STATISTICAL parity against the oracle in MCO_RNG_GLOBAL mode, which is bit-identical to the compiled reference
  (tests/test_oracle_golden.py): k-eff within 3 sigma combined, chi-square over tally bins (north_star).
"""
import numpy as np
import pytest

import golden_cases as gc
import oracle_lib as ol
import mc_old_b200 as mcb
from mc_old_b200 import decks

pytestmark = pytest.mark.gpu


def _close_frac(a, b, rtol=1e-9, atol=0.0):
    a = np.asarray(a); b = np.asarray(b)
    return float(np.mean(np.abs(a - b) <= atol + rtol * np.abs(b)))


# ---------------------------------------------------------------------------------------------------------------
# trajectory parity
# ---------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("mode", ["walk", "split"])
def test_heu_history_parity(mode):
    """HEU sphere, 3 generations of 20000 histories: per-history k_C / k_TL (EstimatorK::end_history,
    Estimator.cpp:514-525), fission-bank size and sites, track / collision counts.  Generations are kept in lock
    step by handing the GPU's bank and the oracle's k to the other side."""
    n = 20000
    deck = mcb.Deck(xml=decks.heu_sphere(samples=n, active=2, passive=1))
    ctx = mcb.Context(deck, device=0, split_stages=(mode == "split"))
    orc = ol.Oracle(deck, rng_mode=ol.RNG_HISTORY, pick_mode=ol.PICK_FLOOR)
    for cyc in range(3):
        g = ctx.run_cycle()
        o, os_, ocell = orc.run_cycle_keep_bank()
        gC, gT = ctx.history_k(n)
        oC, oT = orc.history_k(n)
        assert g.n_histories == n == o.n_histories
        fc, ft = _close_frac(gC, oC, atol=1e-15), _close_frac(gT, oT, atol=1e-15)
        assert fc >= 0.995 and ft >= 0.995, (cyc, fc, ft)
        assert abs(int(g.n_sites) - int(o.n_sites)) <= 0.005 * o.n_sites
        assert abs(int(g.n_tracks) - int(o.n_tracks)) <= 0.005 * o.n_tracks
        assert abs(int(g.n_collisions) - int(o.n_collisions)) <= 0.005 * o.n_collisions
        assert abs(g.k_cycle - o.k_cycle) <= 2e-3 * o.k_cycle
        # the cycle's sums are the exact sums of the per-history scores
        assert g.k_sum_C == pytest.approx(float(np.sum(gC)), rel=1e-12)
        assert g.k_sq_TL == pytest.approx(float(np.sum(gT * gT)), rel=1e-12)
        # fission bank in canonical order: sites of histories that agree are the same sites
        gs, gcell = ctx.fission_bank(int(g.n_sites))
        if gs.shape[0] == os_.shape[0]:
            same = np.all(np.abs(gs[:, :3] - os_[:, :3]) < 1e-9, axis=1)
            assert same.mean() >= 0.98
            assert _close_frac(gs[same, 6], os_[same, 6]) >= 0.995        # Watt energy from the site's own stream
            assert np.mean(np.all(np.abs(gs[same, 3:6] - os_[same, 3:6]) < 1e-9, axis=1)) >= 0.995
        # lock step: same source bank and the same k on both sides for the next generation
        orc.set_source_bank(gs, gcell)
        ctx.k = orc.k
    ctx.close()


def test_heu_entropy_history_parity():
    """per-history Shannon entropy averaged over the cycle (Estimator.cpp:514-528, Entropy.cpp:43-62)"""
    n = 20000
    deck = mcb.Deck(xml=decks.heu_sphere(samples=n, active=1, passive=1, entropy=True))
    ctx = mcb.Context(deck, device=0)
    orc = ol.Oracle(deck, rng_mode=ol.RNG_HISTORY, pick_mode=ol.PICK_FLOOR)
    for cyc in range(2):
        g = ctx.run_cycle(); o = orc.run_cycle()
        assert g.H == pytest.approx(o.H, rel=5e-3), cyc
        assert g.H_cycle_conventional > g.H
        gs, gcell = ctx.fission_bank(int(g.n_sites))
        orc.set_source_bank(gs, gcell); ctx.k = orc.k
    ctx.close()


@pytest.mark.parametrize("name,n", [("slab", 50000), ("slab_overlap", 50000), ("shield", 20000), ("shield_split", 10000), ("fsf", 20000), ("fsf_comb", 20000),
                                    ("heu_tallies", 10000), ("gcr_trmm", 400), ("gcr_trmm_100", 150), ("leak_time", 20000), ("gcr_td", 3000),
                                    ("gcr_td_comb", 3000), ("sphere_det", 200000)])
def test_tallies_history_parity(name, n):
    """Estimator::score / end_history / end_cycle / end_simulation (Estimator.cpp:298-367) on every deck family:
    surface, cell-TL and cell-C estimators, energy filters, splitting (cell_importance, population_control.cpp:21-49)
    and same-history fission secondaries (fixed_source.cpp:12-22); the TRMM tally set of infinite_GCR_TRMM (3500 bins:
    energy_initial x energy matrices filled by simulate-then-score estimators, Estimator.cpp:441-482, delayed-neutron
    scores at energy_old); slab_overlap: a void cell overlapping the one-surface outside cell and preceding it in the deck (search_cell's
    first-match rule against the crossing shortcut of ev_cross_pre: every leaked particle must still reach the far plane), also with the 100 groups of infinite_GCR_TRMM_100 (81500 tallies: more than a grid dimension holds,
    and the keyed form of the per-history tally tables); the <disk_z> source and the cylinder / sphere / plane cells of examples/sphere_detection; time filters that split a track over the bins it spans (examples/HEU_sphere_leakage,
    Estimator.cpp:199-246); the particle comb (population_control.cpp:55-84, fsf_comb: banks of 3 or more waiting particles
    combed to 2); the time-dependent mode (examples/infinite_GCR_TD, _TD_sub: census stops, forced decay of delayed
    neutrons, the <tdmc/> filter; general.cpp:187-195, time_dependent.cpp, fixed_source.cpp:25-40).  Same per-history streams on both sides, so the
    per-bin means agree far inside their statistical error: |gpu - oracle| <= 0.2 sigma + 1e-9 relative."""
    xml = {"slab": lambda: decks.slab(samples=n), "slab_overlap": lambda: decks.slab_overlap(samples=n), "shield": lambda: decks.shielding(samples=n),
           "shield_split": lambda: decks.shielding(samples=n, split=True), "fsf": lambda: decks.fixed_source_fissile(samples=n),
           "fsf_comb": lambda: decks.fixed_source_fissile(samples=n, comb=(3, 2)),
           "heu_tallies": lambda: decks.heu_sphere(samples=n, active=1, passive=0, estimators=True),
           "gcr_trmm": lambda: decks.gcr(samples=n, active=1, passive=0, trmm=True),
           "gcr_trmm_100": lambda: decks.gcr(samples=n, active=1, passive=0, trmm=True, groups=100),
           "leak_time": lambda: decks.heu_leakage(samples=n), "sphere_det": lambda: decks.sphere_detection(samples=n), "gcr_td": lambda: decks.gcr_td(samples=n),
           "gcr_td_comb": lambda: decks.gcr_td(samples=n, linear="1e-8 6 2e-5", comb=(8, 4), groups=10)}[name]()
    deck = mcb.Deck(xml=xml)
    ctx = mcb.Context(deck, device=0)
    orc = ol.Oracle(deck, rng_mode=ol.RNG_HISTORY, pick_mode=ol.PICK_FLOOR)
    g = ctx.run_cycle(); o = orc.run_cycle()
    orc.end_simulation()
    gm, gu = ctx.tallies(); om, ou = orc.tallies()
    if name == "slab_overlap":
        assert gm[0] > 0 and gm[1] == gm[0]
    assert abs(int(g.n_tracks) - int(o.n_tracks)) <= 0.01 * o.n_tracks
    scored = ou > 0
    assert scored.any()
    if name.startswith("gcr_td"):
        # FilterTDMC scores a particle only if its time EQUALS the census time (Estimator.cpp:254) after t += ((T - t) v) / v,
        # which holds or not by the last bit of t: the few histories whose flight lengths differ in the last place between
        # the GPU's log / cos and the host library's (the same 0.5 % test_heu_history_parity allows) score or miss whole
        # particles at a census.  The bins that the filter does not guard agree as everywhere else; the guarded ones to
        # well inside the statistical error, and their sum over all bins of a census to 2 %
        plain = np.arange(len(gm)) >= deck.estimators()[0]["n_tallies"]
        assert np.all(np.abs(gm - om)[plain] <= 1e-9 * np.abs(om[plain]))
        assert np.all(np.abs(gm - om) <= 1.0 * ou + 1e-9 * np.abs(om)), (name, gm, om, ou)
        nt = deck.info["n_tallies"] - int(plain.sum())
        per_census = lambda a: a[:nt].reshape(_n_census(deck), -1).sum(axis=1)
        assert np.allclose(per_census(gm), per_census(om), rtol=0.02)
        assert np.all(np.abs(gu[scored] - ou[scored]) <= 0.25 * ou[scored])
    else:
        assert np.all(np.abs(gm - om) <= 0.2 * ou + 1e-9 * np.abs(om)), (name, gm, om, ou)
        assert np.all(np.abs(gu[scored] - ou[scored]) <= 0.05 * ou[scored])
    ctx.close()


def _n_census(deck):
    return deck.estimators()[0]["filters"][0]["size"]  # the <tdmc/> filter comes first (setup.cpp:703-708)


# ---------------------------------------------------------------------------------------------------------------
# statistical parity with the reference's own stream layout
# ---------------------------------------------------------------------------------------------------------------
def _k_sigma(res, n_passive):
    """standard error of k_avg.  The reference's own k_uncer = sqrt(sum(var_C + var_TL))/Navg/2 (Estimator.cpp:539-549)
    treats the collision and track-length estimators as independent; they are almost fully correlated, so it
    understates the spread of k_avg by ~sqrt(2) (measured here: 8 reference-identical oracle runs of 1e4 x 190 give a
    spread of the cycle means / sqrt(n) of 0.00052 against a reported 0.00034).  Use the larger of the two."""
    k = np.array([r.k_cycle for r in res[n_passive:]])
    return max(res[-1].k_uncer, float(k.std(ddof=1) / np.sqrt(k.size)))


def test_heu_keff_3sigma():
    """k-eff within 3 sigma combined of the reference (north_star).  Reference side: the oracle in GLOBAL mode,
    bit-identical to MC_ref, 1e4 x (10 + 30) generations; GPU side: 1e5 x (10 + 30).  Also anchored on the k the
    survey measured with the reference binary (0.926044 +- 0.000352 reported at 1e4 x 200; x1.5 for the
    correlation the reported figure leaves out, see _k_sigma)."""
    deck_o = mcb.Deck(xml=decks.heu_sphere(samples=10000, active=30, passive=10))
    res = ol.Oracle(deck_o, rng_mode=ol.RNG_GLOBAL, pick_mode=ol.PICK_CDF).run()
    k_o, s_o = res[-1].k_avg, _k_sigma(res, 10)
    deck = mcb.Deck(xml=decks.heu_sphere(samples=100000, active=30, passive=10))
    ctx = mcb.Context(deck, device=0)
    rs = [ctx.run_cycle() for _ in range(40)]
    ctx.close()
    r, s_g = rs[-1], _k_sigma(rs, 10)
    assert abs(r.k_avg - k_o) <= 3 * np.hypot(s_g, s_o), (r.k_avg, s_g, k_o, s_o)
    assert abs(r.k_avg - 0.926044) <= 3 * np.hypot(s_g, 1.5 * 0.000352), (r.k_avg, s_g)
    assert r.k_uncer < res[-1].k_uncer


def test_ucube_keff_and_entropy_3sigma():
    """examples/UCube (the one deck with an <entropy> mesh): k and the per-history-averaged H against the
    reference-identical oracle"""
    deck_o = mcb.Deck(xml=decks.ucube(samples=5000, active=10, passive=5))
    res = ol.Oracle(deck_o, rng_mode=ol.RNG_GLOBAL, pick_mode=ol.PICK_CDF).run()
    k_o, s_o = res[-1].k_avg, res[-1].k_uncer
    H_o = np.array([r.H for r in res[5:]])
    deck = mcb.Deck(xml=decks.ucube(samples=50000, active=10, passive=5))
    ctx = mcb.Context(deck, device=0)
    rs = [ctx.run_cycle() for _ in range(15)]
    ctx.close()
    assert abs(rs[-1].k_avg - k_o) <= 3 * np.hypot(rs[-1].k_uncer, s_o)
    H_g = np.array([r.H for r in rs[5:]])
    sd = np.hypot(H_o.std(ddof=1) / np.sqrt(H_o.size), H_g.std(ddof=1) / np.sqrt(H_g.size))
    assert abs(H_g.mean() - H_o.mean()) <= 3 * sd + 1e-3 * H_o.mean()


@pytest.mark.parametrize("name", ["shield", "fsf"])
def test_tallies_chi_square(name):
    """chi-square over tally bins against the reference-identical oracle run (independent streams): below the
    99.9 % quantile of chi2_N.  Fixed-source decks: histories are independent, so the reported
    uncertainties are the true standard errors."""
    from scipy.stats import chi2
    mk = {"shield": lambda n: decks.shielding(samples=n), "fsf": lambda n: decks.fixed_source_fissile(samples=n)}[name]
    n_o, n_g = 20000, 400000
    deck_o = mcb.Deck(xml=mk(n_o))
    orc = ol.Oracle(deck_o, rng_mode=ol.RNG_GLOBAL, pick_mode=ol.PICK_CDF)
    orc.run()
    om, ou = orc.tallies()
    deck = mcb.Deck(xml=mk(n_g))
    ctx = mcb.Context(deck, device=0)
    for _ in range(deck.info["n_cycle"]):
        ctx.run_cycle()
    gm, gu = ctx.tallies()
    ctx.close()
    ok = (ou > 0) & (gu > 0)
    x2 = float(np.sum((gm[ok] - om[ok]) ** 2 / (gu[ok] ** 2 + ou[ok] ** 2)))
    assert x2 < chi2.ppf(0.999, int(ok.sum())), (name, x2, int(ok.sum()))


def test_tallies_chi_square_k_mode():
    """k-eigenvalue tallies (HEU sphere with cell / energy / surface estimators).  Generations are correlated and
    the scores of one estimator are functions of the same tracks, so neither the reported uncertainties nor a
    chi-square over all 31 tallies are calibrated.  Instead: 16 independent replicas (seeds) per side give the
    standard error of every tally empirically; the chi-square runs over the bins of ONE score per estimator (flux in
    7 energy bins, collision flux, leakage); threshold = 99.9 % quantile of chi2_N x 1.25 for the variance being
    estimated from 16 + 16 replicas."""
    from scipy.stats import chi2
    R = 16

    def replicas(n, run):
        out = []
        for seed in range(1, R + 1):
            deck = mcb.Deck(xml=decks.heu_sphere(samples=n, active=4, passive=4, estimators=True))
            deck.set_run(seed=2 * seed + 1)
            out.append(run(deck))
        return np.array(out)

    def run_oracle(deck):
        orc = ol.Oracle(deck, rng_mode=ol.RNG_GLOBAL, pick_mode=ol.PICK_CDF)
        orc.run()
        return orc.tallies()[0]

    def run_gpu(deck):
        ctx = mcb.Context(deck, device=0)
        for _ in range(deck.info["n_cycle"]):
            ctx.run_cycle()
        m = ctx.tallies()[0]
        ctx.close()
        return m

    o = replicas(5000, run_oracle)
    g = replicas(50000, run_gpu)
    bins = list(range(0, 7)) + [28, 30]  # sphere_rates flux x 7 energy bins, sphere_coll flux, leak cross
    om, gm = o.mean(axis=0), g.mean(axis=0)
    ov, gv = o.var(axis=0, ddof=1) / R, g.var(axis=0, ddof=1) / R
    ok = [b for b in bins if ov[b] > 0 and gv[b] > 0]
    assert len(ok) >= 6
    x2 = float(sum((gm[b] - om[b]) ** 2 / (gv[b] + ov[b]) for b in ok))
    assert x2 < 1.25 * chi2.ppf(0.999, len(ok)), (x2, len(ok))
    # every tally, one by one: within 5 empirical standard errors
    nz = (ov > 0) & (gv > 0)
    z = np.abs(gm[nz] - om[nz]) / np.sqrt(gv[nz] + ov[nz])
    assert z.max() < 5.0, z


def test_gcr_trmm_statistical_gate():
    """BASELINE.json config 2 (examples/infinite_GCR_TRMM: 4 nuclides, reflective planes, the nine TRMM estimators =
    3500 bins that report.cpp:53-158 assembles into the transition-rate matrix): k within 3 sigma combined of the
    reference-identical oracle (RNG_GLOBAL / PICK_CDF, bit-identical to the compiled reference), and a chi-square over
    the tally bins.  Reference side: tests/golden/gcr_trmm_global_4000.npz, 4000 x (5 + 10), made by
    tests/golden/make_gcr_trmm_golden.py (the deck's own 400 histories per generation carry a population-size bias a
    2e4-per-generation run resolves: GPU at N = 400 against the N = 400 reference gives chi2 / dof 1.05 - 1.25, at
    N = 2e4 against the same 2.2 - 2.5).  GPU: 2e4 x (5 + 10), independent streams.

    What the chi-square runs over.  In a k-eigenvalue run all bins of an estimator share the generation-to-generation
    fluctuation of the source (two reference-identical runs differ by ~2 % in EVERY bin at once), and the twelve
    NuFissionDelayed* scores of TRM_simple are one score times constants; a chi-square over the raw 3500 bins counts that
    one fluctuation hundreds of times.  So each (estimator, score) vector is compared as a SHAPE (bins divided by their
    sum), one representative score per family.  Calibration with the oracle alone (two GLOBAL runs of different seeds,
    and GLOBAL against per-history streams, all at N = 400): chi2 / dof = 1.10 ... 1.26 over ~1100 bins — the reported
    per-bin uncertainties leave out the correlation between generations — hence the bound 1.5."""
    import os
    gold = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "gcr_trmm_global_4000.npz"))
    om, ou = gold["mean"], gold["uncer"]
    ko = gold["k_cycle"][5:]
    k_o, s_o = float(gold["k_avg"]), max(float(gold["k_uncer"]), float(ko.std(ddof=1) / np.sqrt(ko.size)))
    deck = mcb.Deck(xml=decks.gcr(samples=20000, active=10, passive=5, trmm=True))
    ctx = mcb.Context(deck, device=0)
    rs = [ctx.run_cycle() for _ in range(15)]
    gm, gu = ctx.tallies()
    ctx.close()
    s_g = _k_sigma(rs, 5)
    assert abs(rs[-1].k_avg - k_o) <= 3 * np.hypot(s_g, s_o), (rs[-1].k_avg, s_g, k_o, s_o)
    assert gm.size == om.size == 3500
    x2, dof = 0.0, 0
    for e in deck.estimators():
        per = e["n_tallies"] // len(e["scores"])
        for k, sc in enumerate(e["scores"]):
            if e["name"] == "TRM_simple" and sc not in ("flux", "NuFissionDelayed_1", "inverse_speed"):
                continue
            sl = slice(e["tally_begin"] + k * per, e["tally_begin"] + (k + 1) * per)
            ok = (ou[sl] > 0) & (gu[sl] > 0)
            if ok.sum() < 2:
                continue
            so, sg = om[sl][ok].sum(), gm[sl][ok].sum()
            x2 += float(np.sum((gm[sl][ok] / sg - om[sl][ok] / so) ** 2 / ((gu[sl][ok] / sg) ** 2 + (ou[sl][ok] / so) ** 2)))
            dof += int(ok.sum()) - 1
            # the totals themselves: within 10 % (the oracle's 4000 active histories fix them to ~2 %)
            assert abs(sg - so) <= 0.10 * so, (e["name"], sc, sg, so)
    assert dof > 900
    assert x2 < 1.5 * dof, (x2, dof)
    # bins the oracle's 4000 active histories never reached carry little weight on the GPU side as well
    rare = (ou == 0) & (gm > 0)
    assert gm[rare].sum() <= 0.02 * gm.sum()


def test_keff_of_the_scaled_up_generation_size():
    """north_star target: HEU_sphere_criticality at 1e8 histories per generation with k-eff within 3 sigma of the
    reference.  The reference cannot run that size; its k is pinned by an ensemble of eight
    reference-identical runs at the deck's size (tests/golden/k_ensemble.json, made by tests/golden/make_k_ensemble.py:
    8 seeds x 1e4 x (10 + 190)).  GPU: 1e8 histories per generation x (12 + 8) generations on one GPU (12 passive: the point source at 14 MeV needs ten generations to settle), in bank batches."""
    import json
    import os
    ens = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "k_ensemble.json")))
    k_ref, s_ref = ens["ensemble_mean"], ens["ensemble_std_of_mean"]
    deck = mcb.Deck(xml=decks.heu_sphere(samples=100_000_000, active=8, passive=12))
    ctx = mcb.Context(deck, device=0)
    rs = [ctx.run_cycle() for _ in range(20)]
    ctx.close()
    assert rs[-1].n_histories == 100_000_000
    s_g = _k_sigma(rs, 12)
    assert abs(rs[-1].k_avg - k_ref) <= 3 * np.hypot(s_g, s_ref), (rs[-1].k_avg, s_g, k_ref, s_ref)
    assert s_g < 1e-4 < s_ref   # the comparison is limited by what the reference can afford, not by the GPU run


def test_slab_analytic_1e7():
    """test/test_integral_Simulator.cpp:20-29: leakage = exp(-(1.2*1 + 0.75*4)); 3 sigma at 1e7 histories"""
    deck = mcb.Deck(xml=decks.slab(samples=10_000_000))
    ctx = mcb.Context(deck, device=0)
    r = ctx.run_cycle()
    m, u = ctx.tallies()
    ctx.close()
    assert r.n_histories == 10_000_000
    assert abs(m[0] - np.exp(-4.2)) <= 3 * u[0], (m[0], u[0])
    assert 3.0e-5 < u[0] < 4.5e-5   # sqrt(p(1-p)/N) = 3.84e-5


def test_sphere_detection_mcnp6_1e7():
    """test/test_integral_Simulator.cpp:10-19: examples/sphere_detection at its own 1e7 histories against MCNP6's
    6.9276e-5 absorptions per source particle in the He-3 tube, with the reference's own acceptance (uncer + 1.15e-6)
    widened to 3 sigma; the <disk_z> source is a superset of the reference's loader (include/mcb200.h)"""
    deck = mcb.Deck(xml=decks.sphere_detection(samples=10_000_000))
    ctx = mcb.Context(deck, device=0)
    r = ctx.run_cycle()
    m, u = ctx.tallies()
    ctx.close()
    assert r.n_histories == 10_000_000
    assert abs(m[1] - 6.9276e-5) <= 3 * u[1] + 1.15e-6, (m, u)
    assert 0.8e-6 < u[1] < 1.6e-6   # 1.82e-6 at 4e6 histories (oracle)


def test_disk_source_outside_every_cell_is_a_lost_particle():
    xml = decks.sphere_detection(samples=2000).replace('<disk_z x="-1.0"  y="0.0" z="-5.0" r="2.0"', '<disk_z x="7.5"  y="0.0" z="-5.0" r="2.0"')
    assert 'x="7.5"' in xml                     # the disk now straddles x = 9 at z = -5: beside the moderator, in "middle vacuum"
    deck = mcb.Deck(xml=xml)                    # ... which exists, so this still loads and runs
    ctx = mcb.Context(deck, device=0); ctx.run_cycle(); ctx.close()
    xml = xml.replace('<cell name="middle vacuum" importance="0.0">\n        <surface name="cx1" sense="+1"/>\n        <surface name="px1" sense="+1"/>\n'
                      '        <surface name="px2" sense="-1"/>\n    </cell>', "")
    assert "middle vacuum" not in xml
    ctx = mcb.Context(mcb.Deck(xml=xml), device=0)
    with pytest.raises(RuntimeError, match="particle is lost"):
        ctx.run_cycle()
    ctx.close()


# ---------------------------------------------------------------------------------------------------------------
# size-independent properties at BASELINE.json's full size, determinism, host-buffer round trips, errors
# ---------------------------------------------------------------------------------------------------------------
def test_full_size_generation_properties():
    """HEU sphere at 1e7 histories per generation (BASELINE.json configs[2]): conservation identities of the event
    loop and physical invariants of the fission bank"""
    n = 10_000_000
    deck = mcb.Deck(xml=decks.heu_sphere(samples=n, active=1, passive=1))
    ctx = mcb.Context(deck, device=0)
    for cyc in range(2):
        r = ctx.run_cycle()
        assert r.n_histories == n
        assert r.n_tracks == r.n_collisions + r.n_crossings      # every flight ends in exactly one event
        assert r.n_lookups == r.n_tracks                          # one lookup per flight in a material
        assert 0.8 < r.k_cycle < 3.0
        kC, kT = ctx.history_k(n)
        assert r.k_sum_C == pytest.approx(float(np.sum(kC)), rel=1e-11)
        assert r.k_sum_TL == pytest.approx(float(np.sum(kT)), rel=1e-11)
        assert r.k_cycle == (r.k_sum_C / n + r.k_sum_TL / n) / 2
        sites, cells = ctx.fission_bank(int(r.n_sites))
        assert sites.shape[0] == r.n_sites
        assert np.all(np.sum(sites[:, :3] ** 2, axis=1) <= 7.68 ** 2 * (1 + 1e-12))
        assert np.allclose(np.sum(sites[:, 3:6] ** 2, axis=1), 1.0, atol=1e-12)
        assert np.all(sites[:, 6] > 0) and np.all(cells == 0)
        # E[sites per history] = k_C estimator / k_prev : the implicit-fission banking is unbiased
        assert r.n_sites == pytest.approx(r.k_sum_C / (1.0 if cyc == 0 else k_prev), rel=5e-3)
        k_prev = r.k_cycle
    ctx.close()


def test_batching_does_not_change_results():
    """a generation processed in several bank batches (small bank_capacity) gives bit-identical k sums (exact
    fixed-point accumulation) and an identical fission bank (canonical order)"""
    n = 50000
    deck = mcb.Deck(xml=decks.heu_sphere(samples=n, active=1, passive=1))
    out = []
    for cap in (0, 12000, 7001):
        ctx = mcb.Context(deck, device=0, bank_capacity=cap)
        rs = [ctx.run_cycle() for _ in range(2)]
        bank = ctx.fission_bank(int(rs[-1].n_sites))
        out.append((rs, bank))
        ctx.close()
    for rs, bank in out[1:]:
        for a, b in zip(out[0][0], rs):
            assert (a.k_sum_C, a.k_sum_TL, a.k_sq_C, a.k_sq_TL, a.n_sites, a.n_tracks) == \
                   (b.k_sum_C, b.k_sum_TL, b.k_sq_C, b.k_sq_TL, b.n_sites, b.n_tracks)
        assert np.array_equal(out[0][1][0], bank[0]) and np.array_equal(out[0][1][1], bank[1])


def test_event_queue_and_walk_kernels_agree():
    """the event-queue formulation (one kernel per event type) and the history-walk kernel are the same computation"""
    n = 30000
    deck = mcb.Deck(xml=decks.ucube(samples=n, active=1, passive=1))
    res = []
    for split in (False, True):
        ctx = mcb.Context(deck, device=0, split_stages=split)
        rs = [ctx.run_cycle() for _ in range(2)]
        res.append((rs, ctx.fission_bank(int(rs[-1].n_sites))))
        ctx.close()
    for a, b in zip(res[0][0], res[1][0]):
        assert (a.k_sum_C, a.k_sum_TL, a.H, a.n_sites, a.n_tracks, a.n_collisions) == \
               (b.k_sum_C, b.k_sum_TL, b.H, b.n_sites, b.n_tracks, b.n_collisions)
    assert np.array_equal(res[0][1][0], res[1][1][0])


@pytest.mark.parametrize("deck_name", ["ucube", "shield_split"])
def test_walk_forms_agree(deck_name, monkeypatch):
    """the two forms of the walk kernel — history per lane, and particles sorted by next event through the
    shared-memory collide / cross queues (MCB_WALK_EXCHANGE=1) — are the same computation: identical k sums, entropy,
    counts and fission bank; tallies to rounding"""
    xml = decks.ucube(samples=30000, active=1, passive=1) if deck_name == "ucube" else decks.shielding(samples=20000, split=True)
    deck = mcb.Deck(xml=xml)
    res = []
    for exch in ("0", "1"):
        monkeypatch.setenv("MCB_WALK_EXCHANGE", exch)
        ctx = mcb.Context(deck, device=0)
        rs = [ctx.run_cycle() for _ in range(deck.info["n_cycle"] if deck_name != "ucube" else 2)]
        bank = ctx.fission_bank(int(rs[-1].n_sites)) if deck_name == "ucube" else None
        res.append((rs, bank, ctx.tallies()))
        ctx.close()
    for other in res[1:]:
        for a, b in zip(res[0][0], other[0]):
            assert (a.k_sum_C, a.k_sum_TL, a.H, a.n_sites, a.n_tracks, a.n_collisions, a.n_crossings) == \
                   (b.k_sum_C, b.k_sum_TL, b.H, b.n_sites, b.n_tracks, b.n_collisions, b.n_crossings)
        if res[0][1] is not None:
            assert np.array_equal(res[0][1][0], other[1][0])
        assert np.allclose(res[0][2][0], other[2][0], rtol=1e-12, atol=0)


def test_source_bank_host_round_trip():
    """mcb_get_source_bank -> mcb_set_source_bank (the host-buffer form of Sbank = Fbank, handler.cpp:16) is the
    identity, and an empty bank is the reference's "[ERROR] Source bank is empty..." (Source.cpp:33-39)"""
    n = 20000
    deck = mcb.Deck(xml=decks.heu_sphere(samples=n, active=2, passive=1))
    a = mcb.Context(deck, device=0); b = mcb.Context(deck, device=0)
    ra = a.run_cycle(); rb = b.run_cycle()
    sites, cells = a.source_bank(int(ra.n_sites))
    assert sites.shape[0] == ra.n_sites
    b.set_source_bank(sites, cells)
    s2, c2 = b.source_bank(int(ra.n_sites))
    assert np.array_equal(sites, s2) and np.array_equal(cells, c2)
    ra = a.run_cycle(); rb = b.run_cycle()
    assert (ra.k_sum_C, ra.k_sum_TL, ra.n_sites) == (rb.k_sum_C, rb.k_sum_TL, rb.n_sites)
    b.set_source_bank(np.zeros((0, 8)), np.zeros(0, dtype=np.int32))
    with pytest.raises(RuntimeError, match="Source bank is empty"):
        b.run_cycle()
    a.close(); b.close()


def test_host_bank_cycle_equals_the_three_plain_calls():
    """mcb_run_cycle_host (bank in host buffers on both sides, upload pipelined with the walk through sorted draws) is
    bit-identical to mcb_set_source_bank + mcb_run_cycle + mcb_get_source_bank"""
    n = 50000
    deck = mcb.Deck(xml=decks.heu_sphere(samples=n, active=2, passive=1, entropy=True, estimators=True))
    a = mcb.Context(deck, device=0); b = mcb.Context(deck, device=0)
    ra = a.run_cycle(); rb = b.run_cycle()
    sites, cells = a.source_bank(int(ra.n_sites))
    for _ in range(2):  # the two active cycles: tallies are scored on both paths
        a.set_source_bank(sites, cells)
        ra = a.run_cycle()
        sa, ca = a.source_bank(int(ra.n_sites))
        out_s = np.zeros((4 * n, 8)); out_c = np.zeros(4 * n, dtype=np.int32)
        rb, sb, cb = b.run_cycle_host(sites, cells, out_s, out_c)
        assert (ra.k_sum_C, ra.k_sum_TL, ra.k_sq_C, ra.H, ra.n_sites, ra.n_tracks, ra.n_collisions) == \
               (rb.k_sum_C, rb.k_sum_TL, rb.k_sq_C, rb.H, rb.n_sites, rb.n_tracks, rb.n_collisions)
        assert np.array_equal(sa, sb) and np.array_equal(ca, cb)
        sites, cells = sa.copy(), ca.copy()
    # tallies: per-history scores are identical; the cycle sums are accumulated with floating-point reductions whose
    # order is not fixed, so they agree to rounding
    ta, tb = a.tallies(), b.tallies()
    assert np.allclose(ta[0], tb[0], rtol=1e-12, atol=0) and np.allclose(ta[1], tb[1], rtol=1e-9, atol=0) and np.count_nonzero(ta[0]) > 20
    with pytest.raises(RuntimeError, match="Source bank is empty"):
        b.run_cycle_host(np.zeros((0, 8)), np.zeros(0, dtype=np.int32), out_s, out_c)
    a.close(); b.close()


def test_capacity_and_lost_particle_errors():
    """a fission bank that overflows is an error, not a truncation; a particle that leaves every cell is the
    reference's "[WARNING] A particle is lost" (general.cpp:31-33) as a status code"""
    deck = mcb.Deck(xml=decks.heu_sphere(samples=20000, active=1, passive=1))
    ctx = mcb.Context(deck, device=0, site_capacity=1000)
    with pytest.raises(RuntimeError, match="fission bank overflow"):
        ctx.run_cycle()
    ctx.close()
    # a slab whose right-hand outside cell is missing: particles crossing x = 5 find no cell
    xml = decks.slab(samples=1000).replace('<cell name="right outside" importance="0.0">\n        <surface name="px3" sense="+1"/>\n    </cell>', "")
    assert "right outside" not in xml
    deck = mcb.Deck(xml=xml)
    ctx = mcb.Context(deck, device=0)
    with pytest.raises(RuntimeError, match="particle is lost"):
        ctx.run_cycle()
    ctx.close()


def test_reproducible_across_contexts():
    """same deck, same seed -> bit-identical generations (per-history streams, exact sums, canonical bank order)"""
    deck = mcb.Deck(xml=decks.heu_sphere(samples=30000, active=2, passive=1))
    runs = []
    for _ in range(2):
        ctx = mcb.Context(deck, device=0)
        rs = [ctx.run_cycle() for _ in range(3)]
        runs.append(([(r.k_cycle, r.k_sq_C, r.n_sites, r.n_tracks) for r in rs], ctx.fission_bank(int(rs[-1].n_sites))[0]))
        ctx.close()
    assert runs[0][0] == runs[1][0]
    assert np.array_equal(runs[0][1], runs[1][1])


@pytest.mark.parametrize("bank", ["peer_reads", "nccl_gather"])
def test_multi_gpu_matches_single_gpu(bank):
    """2 ranks over NCCL (one process per GPU), with the bank read in place from the peer (CUDA IPC, sorted draws) or
    gathered with NCCL: k, H, track counts and the bank per generation are bit-identical to the 1-GPU run (exact
    integer sums, canonical bank order, per-history streams); tallies (double sums in rank order) agree
    to rounding"""
    import json
    import os
    import subprocess
    import sys
    if mcb.cuda_lib().mcb_device_count() < 2:
        pytest.skip("needs 2 GPUs")
    worker = os.path.join(os.path.dirname(os.path.abspath(__file__)), "mgpu_worker.py")
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                          "--master-addr", "127.0.0.1", "--master-port", "29611", worker, "--samples", "40000", "--cycles", "3"],
                         capture_output=True, text=True, timeout=600,
                         env=dict(os.environ, **({"MCB_NO_P2P": "1"} if bank == "nccl_gather" else {})))
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    multi = json.loads([l for l in out.stdout.splitlines() if l.startswith("{")][-1])
    deck = mcb.Deck(xml=decks.heu_sphere(samples=40000, active=2, passive=1, entropy=True, estimators=True))
    ctx = mcb.Context(deck, device=0)
    rs = [ctx.run_cycle() for _ in range(3)]
    sites, cells = ctx.source_bank(int(rs[-1].n_sites))
    mean, uncer = ctx.tallies()
    bank_energy_sum = float(np.sum(sites[:, 6])).hex()
    hk = []
    for _ in range(2):  # host-bank cycles: on several GPUs every rank moves only its slice of the bank (peer reads path)
        out_s = np.zeros((4 * 40000, 8)); out_c = np.zeros(4 * 40000, dtype=np.int32)
        r, s_new, c_new = ctx.run_cycle_host(sites, cells, out_s, out_c)
        sites, cells = s_new.copy(), c_new.copy()
        hk.append([r.k_cycle.hex(), int(r.n_sites), int(r.n_tracks)])
    ctx.close()
    assert hk == multi["host_cycles"]
    assert [float(np.sum(sites[:, i])).hex() for i in range(8)] == multi["host_bank_sum_hex"]
    assert [r.H.hex() for r in rs] == multi["H_hex"]
    assert [int(r.n_tracks) for r in rs] == multi["n_tracks"]
    assert np.allclose(mean, multi["tally_mean"], rtol=1e-12, atol=0) and np.allclose(uncer, multi["tally_uncer"], rtol=1e-9, atol=0)
    assert [r.k_cycle.hex() for r in rs] == multi["k_cycle_hex"]
    assert [int(r.n_sites) for r in rs] == multi["n_sites"]
    assert bank_energy_sum == multi["bank_energy_sum_hex"]
