"""Function-level parity of the CUDA path, called through the C-ABI (include/mcb200.h), against
(i) the committed golden vectors of the compiled reference (tests/golden/functions.npz) and
(ii) the oracle (oracle/mc_oracle.c) on larger seeded inputs.

Bar (north_star): cross-section lookups and reaction-channel selection are BIT-EXACT for the same energy and
random inputs; so are the RNG streams, the geometry (IEEE sqrt/div only, no FMA contraction on either side) and
the cell search.  The sampled kinematics (free-gas scatter, Watt) go through libm (log, cos, sin), which differs
from glibc at the ulp level on the GPU: tolerance 1e-9 relative, stated per test.
"""
import ctypes as C

import numpy as np
import pytest

import golden_cases as gc
import oracle_lib as ol
import mc_old_b200 as mcb

pytestmark = pytest.mark.gpu

FN_DECKS = gc.function_decks()


@pytest.fixture(scope="module")
def ctxs(deck_cache):
    cache = {}

    def get(name):
        if name not in cache:
            deck = deck_cache(FN_DECKS[name])
            cache[name] = (deck, mcb.Context(deck, device=0))
        return cache[name]
    yield get
    for _, c in cache.values():
        c.close()


def _mats(golden_fn, name):
    return sorted({int(k.split("/")[2]) for k in golden_fn.files if k.startswith("xs/%s/" % name)})


def _nucs(golden_fn, name):
    return sorted({int(k.split("/")[2]) for k in golden_fn.files if k.startswith("micro/%s/" % name)})


@pytest.mark.parametrize("name", sorted(FN_DECKS))
def test_xs_lookup_golden_bit_exact(name, golden_fn, ctxs):
    """Material::Sigma{T,S,C,F}, nuSigmaF (Material.cpp:18-65): grid points +-1 ulp, duplicate energies, clamps"""
    deck, ctx = ctxs(name)
    for m in _mats(golden_fn, name):
        E = golden_fn["xs/%s/%d/E" % (name, m)]
        want = golden_fn["xs/%s/%d/sigma" % (name, m)]
        got = ctx.xs_lookup(m, E)
        assert np.array_equal(got, want[:, :5]), "material %d of %s" % (m, name)


@pytest.mark.parametrize("name", sorted(FN_DECKS))
def test_xs_lookup_vs_oracle_1e6(name, ctxs):
    """SURVEY App. J gate: >= 1e6 log-uniform energies in [1e-6, 3e8] eV per material, memcmp-equal to the oracle"""
    deck, ctx = ctxs(name)
    r = np.random.default_rng(12345)
    for m in range(deck.info["n_materials"]):
        E = np.exp(r.uniform(np.log(1e-6), np.log(3e8), 1_000_000))
        got = ctx.xs_lookup(m, E)
        want = ol.xs_lookup(deck, m, E)
        assert got.tobytes() == want.tobytes(), "material %d of %s" % (m, name)


@pytest.mark.parametrize("name", sorted(FN_DECKS))
def test_xs_lookup_every_grid_point(name, ctxs):
    """every grid energy of every nuclide of every material and its +-1 ulp neighbours (union-grid / hash edges,
    duplicate energies, and the two descents of the B-11 grid that the reference bisects as they are)"""
    deck, ctx = ctxs(name)
    for m, zaids in enumerate(gc._MAT_ZAIDS[name]):
        g = np.concatenate([gc.grid_energies(z) for z in zaids])
        E = np.concatenate([g, np.nextafter(g, 0.0), np.nextafter(g, np.inf)])
        assert ctx.xs_lookup(m, E).tobytes() == ol.xs_lookup(deck, m, E).tobytes(), "material %d of %s" % (m, name)


def test_xs_lookup_edge_inputs(ctxs):
    """empty batch, one element, E = 0 (quirk 11 is a memo artefact the stateless lookup does not have: E=0 clamps
    low like any E below the grid), denormal, huge"""
    deck, ctx = ctxs("heu")
    assert ctx.xs_lookup(0, np.zeros(0)).shape == (0, 5)
    E = np.array([5e-324, 1e-300, 1e-5, 2e7, 1e30, 1.7e308])
    assert ctx.xs_lookup(0, E).tobytes() == ol.xs_lookup(deck, 0, E).tobytes()
    assert ctx.xs_lookup(0, E[:1]).tobytes() == ol.xs_lookup(deck, 0, E[:1]).tobytes()
    with pytest.raises(RuntimeError):
        ctx.xs_lookup(7, E)


def test_xs_spot_values(ctxs):
    """oracle spot values of the HEU material printed with %.17g"""
    _, ctx = ctxs("heu")
    got = ctx.xs_lookup(0, np.array([0.0253, 1e6, 1.4e7]))
    want = np.array([[31.842638548511541, 0.71937063933950007, 4.5002489565894006, 26.62301895258264, 64.872310284905424],
                     [0.24198625830147802, 0.18202392938330003, 0.0053564364232400003, 0.05460589249493801, 0.13830207652539755],
                     [0.23000679985844003, 0.13142263971280002, 5.8167856239999998e-05, 0.098525992289399997, 0.43259314430759871]])
    assert np.array_equal(got, want)


@pytest.mark.parametrize("name", sorted(FN_DECKS))
def test_select_channel_bit_exact(name, golden_fn, ctxs):
    """Material::nuclide_scatter / nuclide_nufission (Material.cpp:106-125) with injected xi: golden + 2e5 vs oracle"""
    deck, ctx = ctxs(name)
    r = np.random.default_rng(777)
    for m in _mats(golden_fn, name):
        for kind in (0, 1):
            key = "select/%s/%d/%d" % (name, m, kind)
            got = ctx.select_channel(m, kind, golden_fn[key + "/E"], golden_fn[key + "/xi"])
            assert np.array_equal(got, golden_fn[key + "/out"]), key
            E = np.exp(r.uniform(np.log(1e-6), np.log(3e8), 200_000))
            xi = r.uniform(0.0, 1.0, E.size)
            xi[:4] = [0.0, 1.0 - 2.0 ** -53, 2.0 ** -63, 0.5]
            assert np.array_equal(ctx.select_channel(m, kind, E, xi), ol.select_channel(deck, m, kind, E, xi)), key


@pytest.mark.parametrize("name", sorted(FN_DECKS))
def test_beta_bit_exact(name, golden_fn, ctxs):
    """Nuclide::beta (Nuclide.cpp:74-77) of every nuclide that sits in a material"""
    deck, ctx = ctxs(name)
    P = _problem(deck)
    for m in range(P.n_materials):
        for ln in range(P.mat_begin[m + 1] - P.mat_begin[m]):
            n = P.mat_nuclide[P.mat_begin[m] + ln]
            key = "micro/%s/%d" % (name, n)
            if key + "/E" not in golden_fn.files:
                continue
            E = golden_fn[key + "/E"]
            assert np.array_equal(ctx.beta(m, ln, E), golden_fn[key + "/out"][:, 5]), key


class _Problem(C.Structure):  # the leading fields of mcb_problem that the tests read
    _fields_ = [("abi_version", C.c_int32), ("ksearch", C.c_int32), ("n_sample", C.c_uint64), ("n_cycle", C.c_uint64),
                ("n_passive", C.c_uint64), ("wr", C.c_double), ("ws", C.c_double), ("seed", C.c_uint64),
                ("n_nuclides", C.c_int32), ("n_materials", C.c_int32), ("nuclides", C.c_void_p),
                ("xs_rows", C.c_void_p), ("n_xs_rows", C.c_int64), ("delayed_data", C.c_void_p),
                ("n_delayed_data", C.c_int64), ("mat_begin", C.POINTER(C.c_int32)),
                ("mat_nuclide", C.POINTER(C.c_int32)), ("mat_density", C.POINTER(C.c_double))]


def _problem(deck):
    return _Problem.from_address(deck.problem)


def test_rng_streams_bit_exact(ctxs):
    """Urand / RN_skip_ahead / RN_init_particle (Random.cpp:121-149,196-204): RN_CHECK known answers through the
    GPU's skip-ahead, and per-history streams equal to the oracle's"""
    _, ctx = ctxs("heu")
    check = [3512401965023503517, 5461769869401032777, 1468184805722937541, 5160872062372652241, 6637647758174943277]
    got = ctx.rng(np.array([0], dtype=np.uint64), 5)
    assert got[0].tolist() == check
    L = ol.oracle()
    nps = np.array([0, 1, 2, 12345, 10 ** 9, 2 ** 40 + 7, 2 ** 62], dtype=np.uint64)
    got = ctx.rng(nps, 3)
    for i, n in enumerate(nps.tolist()):
        s = L.mco_lcg_skip(1, (n * 152917) & (2 ** 64 - 1))
        for d in range(3):
            s = L.mco_lcg_next(s)
            assert int(got[i, d]) == s
    # stream 123456/152917 is not an integer: check the raw skip through draws of history 0 instead
    long = ctx.rng(np.array([0], dtype=np.uint64), 123460)[0]
    assert long[123455:].tolist() == [794206257475890433, 4662153896835267997, 6075201270501039433,
                                     889694366662031813, 7299299962545529297]


@pytest.mark.parametrize("name", sorted(FN_DECKS))
def test_geometry_bit_exact(name, golden_fn, ctxs):
    """surface_intersect over each cell's surfaces (general.cpp:54-67) and Surface::eval/distance
    (Geometry.cpp:29-188): distance, surface and eval equal the oracle's bit for bit; the golden per-surface
    distances of the reference are the minimum's candidates"""
    deck, ctx = ctxs(name)
    L = ol.oracle()
    pos, dirs, want = (golden_fn["surface/%s/%s" % (name, k)] for k in ("pos", "dir", "out"))
    ncell = deck.info["n_cells"]
    r = np.random.default_rng(5)
    big_pos = np.concatenate([pos, r.uniform(-8, 10, (20000, 3))])
    d = r.normal(size=(20000, 3)); d /= np.linalg.norm(d, axis=1)[:, None]
    big_dir = np.concatenate([dirs, d])
    for c in range(ncell):
        cells = np.full(big_pos.shape[0], c, dtype=np.int32)
        got = ctx.geometry(cells, big_pos, big_dir)
        for i in list(range(pos.shape[0])) + list(range(pos.shape[0], big_pos.shape[0], 97)):
            p = np.ascontiguousarray(big_pos[i]); u = np.ascontiguousarray(big_dir[i])
            dist = C.c_double()
            S = L.mco_surface_intersect(deck.problem, c, p.ctypes.data, u.ctypes.data, C.byref(dist))
            assert got[i, 0] == dist.value and int(got[i, 1]) == S, (name, c, i)
            if S >= 0:
                assert got[i, 2] == L.mco_surface_eval(deck.problem, S, p.ctypes.data)
                if i < pos.shape[0]:
                    assert got[i, 0] == want[S, i, 1] and got[i, 2] == want[S, i, 0]


@pytest.mark.parametrize("name", sorted(FN_DECKS))
def test_search_cell_bit_exact(name, ctxs):
    """search_cell / test_point (general.cpp:13-34): first cell in deck order, -1 when lost"""
    deck, ctx = ctxs(name)
    L = ol.oracle()
    r = np.random.default_rng(9)
    pos = r.uniform(-12, 12, (5000, 3))
    got = ctx.search_cell(pos)
    for i in range(pos.shape[0]):
        p = np.ascontiguousarray(pos[i])
        assert got[i] == L.mco_search_cell(deck.problem, p.ctypes.data)
        assert got[i] == deck.search_cell(*pos[i])


def test_scatter_and_watt_samples(ctxs, golden_fn):
    """ReactionScatter::sample (Reaction.cpp:27-118) and DistributionWatt::sample (Distribution.cpp:34-73) from the
    stream of history nps: the GPU draws the same xi (integer LCG) and evaluates log/cos/sin with CUDA's libm.
    Tolerance: 1e-9 relative on energy/speed, 1e-9 absolute on direction cosines, for >= 99.9 % of the samples
    (a rejection test decided the other way by an ulp changes the whole sample)."""
    for name in ("heu", "gcr", "shield"):
        deck, ctx = ctxs(name)
        L = ol.oracle()
        r = np.random.default_rng(31)
        n = 20000
        nps = r.integers(0, 2 ** 40, n, dtype=np.uint64)
        E = np.exp(r.uniform(np.log(1e-3), np.log(2e7), n))
        d = r.normal(size=(n, 3)); d /= np.linalg.norm(d, axis=1)[:, None]
        for nuc in range(deck.info["n_nuclides"]):
            io = np.concatenate([d, E[:, None], np.zeros((n, 1))], axis=1)
            got = ctx.scatter(nuc, nps, io)
            want = np.zeros_like(io)
            for i in range(0, n, 10):
                s = C.c_uint64(L.mco_lcg_skip(1, (int(nps[i]) * 152917) & (2 ** 64 - 1)))
                w = io[i].copy()
                L.mco_scatter_sample(deck.problem, nuc, C.byref(s), w.ctypes.data)
                want[i] = w
            sel = slice(0, n, 10)
            ok = np.all(np.abs(got[sel, :3] - want[sel, :3]) < 1e-9, axis=1) & \
                np.all(np.abs(got[sel, 3:] - want[sel, 3:]) <= 1e-9 * np.abs(want[sel, 3:]), axis=1)
            assert ok.mean() >= 0.999, (name, nuc, ok.mean())
            assert np.allclose(np.linalg.norm(got[:, :3], axis=1), 1.0, atol=1e-12)
        # Watt spectrum of the fissile nuclides
        fissile = golden_fn["watt/%s/seed" % name].any(axis=1)  # the reference's Watt loop never ends for a = b = 0
        for nuc in range(deck.info["n_nuclides"]):
            if not fissile[nuc]:
                continue
            got = ctx.watt(nuc, nps, E)
            want = []
            for i in range(0, n, 10):
                s = C.c_uint64(L.mco_lcg_skip(1, (int(nps[i]) * 152917) & (2 ** 64 - 1)))
                want.append(L.mco_watt_sample(deck.problem, nuc, C.byref(s), float(E[i])))
            want = np.array(want)
            ok = np.abs(got[::10] - want) <= 1e-9 * np.abs(want)
            assert ok.mean() >= 0.999, (name, nuc, ok.mean())


def test_watt_spectrum_mean(ctxs):
    """size-independent property: mean of the Watt spectrum = a*(3/2 + a*b/4) MeV (U-235 thermal: a=0.988, b=2.249)"""
    _, ctx = ctxs("heu")
    n = 2_000_000
    got = ctx.watt(0, np.arange(n, dtype=np.uint64), np.full(n, 0.0253))
    a, b = 0.988, 2.249
    mean = a * (1.5 + a * b / 4.0) * 1e6
    sd = got.std() / np.sqrt(n)
    assert abs(got.mean() - mean) < 4 * sd


def test_shared_reciprocal_division_is_ieee(ctxs):
    """csrc/mcb_physics.h: quotients of one denominator share a refined reciprocal (interpolation weights of
    XSTable::xs, vector normalisations of the scatter kinematics).  That form must be the IEEE division bit for bit,
    or cross sections stop being bit-exact against the reference's x86 divisions: 4e6 random pairs over the whole
    exponent range the path meets, grid-like nearly equal operands, exact quotients, zeros, denormals, infinities;
    the GPU's plain division is in turn checked against numpy's."""
    _, ctx = ctxs(sorted(FN_DECKS)[0])
    rng = np.random.default_rng(20261018)
    n = 1_000_000
    a = np.concatenate([
        rng.standard_normal(n) * 10.0 ** rng.uniform(-30, 30, n),
        rng.uniform(1e-5, 2e7, n),                                   # energies
        rng.uniform(-1, 1, n),                                       # direction components
        (rng.integers(1, 1 << 20, n) * rng.integers(1, 1 << 20, n)).astype(np.float64),  # exact quotients
        np.array([0.0, -0.0, 1e-310, 5e-324, 1e308, np.inf, 1.0, 3.0, 1e-300, 2.0 ** -969, 2.0 ** -970]),
    ])
    b = np.concatenate([
        rng.standard_normal(n) * 10.0 ** rng.uniform(-30, 30, n),
        rng.uniform(1e-5, 2e7, n) * (1 + rng.uniform(-1e-12, 1e-12, n)),
        rng.uniform(1e-3, 1e9, n),
        rng.integers(1, 1 << 20, n).astype(np.float64),
        np.array([3.0, 7.0, 3.0, 3.0, 1e-10, 2.0, 1e-310, 0.0, 1e300, 3.0, 3.0]),
    ])
    b[b == 0.0] = 1.5
    b[-4] = 0.0
    shared, plain = ctx.division(a, b)
    bad = np.nonzero((shared.view(np.uint64) != plain.view(np.uint64)) & ~(np.isnan(shared) & np.isnan(plain)))[0]
    assert bad.size == 0, [(float(a[i]).hex(), float(b[i]).hex(), float(shared[i]).hex(), float(plain[i]).hex()) for i in bad[:8]]
    with np.errstate(all="ignore"):
        want = a / b
    assert np.array_equal(plain.view(np.uint64)[~np.isnan(want)], want.view(np.uint64)[~np.isnan(want)])
