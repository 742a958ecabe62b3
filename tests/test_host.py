"""Host-side logic, CPU only: the C-ABI libraries load and export every symbol include/*.h declares, the deck
loader (replacement of the reference's Simulator constructor, setup.cpp:30-1069) keeps the grammar, defaults and
error messages, the history sharding rule, and the N > 1 data flow (shard -> transport -> all-reduce of the sums +
all-gather of the fission bank -> identical close-out on every rank) over gloo with world_size 2.
No compute call touches a GPU here.
"""
import ctypes as C
import os
import re
import sys

import numpy as np
import pytest

import golden_cases as gc
import oracle_lib as ol
import mc_old_b200 as mcb
from mc_old_b200 import decks

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared(header, prefix):
    text = open(os.path.join(ROOT, "include", header)).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(%s_[a-z0-9_]+)\s*\(" % prefix, text)))


def test_cuda_library_exports_every_declared_symbol():
    """include/mcb200.h is the drop-in boundary: every entry point it declares must be in libmcb200.so"""
    names = _declared("mcb200.h", "mcb")
    assert len(names) >= 25
    L = C.CDLL(mcb.CUDA_LIB)
    for n in names:
        assert hasattr(L, n), n


def test_host_library_exports_every_declared_symbol():
    names = _declared("mcb200_host.h", "mcbh")
    assert len(names) >= 12
    L = C.CDLL(mcb.HOST_LIB)
    for n in names:
        assert hasattr(L, n), n


def test_no_cpu_fallback():
    """without a device the product refuses to compute (no CPU fallback)"""
    if mcb.cuda_lib().mcb_device_count() > 0:
        pytest.skip("a GPU is present")
    deck = mcb.Deck(xml=decks.slab(samples=10))
    with pytest.raises(RuntimeError, match="no CUDA device"):
        mcb.Context(deck)


def test_product_does_not_touch_the_oracle():
    """nothing under mc_old_b200/ or bench.py's own arm links, imports or opens oracle code"""
    for dirpath, _, files in os.walk(os.path.join(ROOT, "mc_old_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                text = open(os.path.join(dirpath, f)).read()
                assert "mc_oracle" not in text and "oracle_lib" not in text and "libref_harness" not in text, f
                for line in text.splitlines():  # no path into oracle/ either; comments may mention the oracle patches
                    code = line.split("//")[0].split("#")[0]
                    assert "oracle" not in code or "CPU oracle" in code, (f, line)


@pytest.mark.parametrize("n,world", [(10, 1), (10, 3), (7, 8), (10 ** 9, 8), (2 ** 40 + 3, 7), (0, 4)])
def test_shard_range_partitions_the_histories(n, world):
    """contiguous rank-ordered slices that cover [0, n) exactly"""
    end = 0
    sizes = []
    for r in range(world):
        b, c = mcb.shard_range(n, r, world)
        assert b == end
        end = b + c
        sizes.append(c)
    assert end == n
    assert max(sizes) - min(sizes) <= 1


# ---- deck loader -------------------------------------------------------------------------------------------------
def test_deck_facts_of_the_baseline_configs():
    """sizes of BASELINE.json's decks as SURVEY §8 lists them"""
    d = mcb.Deck(xml=decks.slab(samples=1e6)).info
    assert (d["n_sample"], d["ksearch"], d["n_nuclides"], d["n_materials"], d["n_surfaces"], d["n_cells"], d["n_tallies"]) == \
           (10 ** 6, 0, 3, 2, 3, 4, 1)
    d = mcb.Deck(xml=decks.heu_sphere(samples=1e7, active=150, passive=50)).info
    assert (d["n_sample"], d["n_cycle"], d["n_passive"], d["ksearch"], d["n_nuclides"], d["n_surfaces"], d["n_cells"]) == \
           (10 ** 7, 200, 50, 1, 2, 1, 2)
    assert d["n_xs_rows"] == 41010 + 162680
    d = mcb.Deck(xml=decks.shielding(samples=1e8)).info
    assert (d["n_nuclides"], d["n_materials"], d["n_surfaces"], d["n_cells"], d["n_estimators"], d["n_tallies"]) == (6, 3, 8, 10, 1, 2)
    d = mcb.Deck(xml=decks.gcr(samples=400, active=100, passive=10)).info
    assert (d["n_nuclides"], d["n_materials"], d["n_surfaces"], d["n_cells"]) == (4, 1, 2, 1)
    assert d["n_xs_rows"] == 41010 + 162680 + 2714 + 1021
    d = mcb.Deck(xml=decks.ucube(samples=1e4)).info
    assert d["entropy_on"] == 1


def test_deck_source_forms():
    """<point x y z .../> (setup.cpp:1051-1063) and the HEU example's own <source position= .../> form, which the
    reference rejects and this loader accepts as a superset"""
    a = mcb.Deck(xml=decks.heu_sphere(samples=100, source_tag="point"))
    b = mcb.Deck(xml=decks.heu_sphere(samples=100, source_tag="source"))
    assert a.info["n_sources"] == b.info["n_sources"] == 1
    o1 = ol.Oracle(a, rng_mode=ol.RNG_GLOBAL, pick_mode=ol.PICK_CDF).run_cycle()
    o2 = ol.Oracle(b, rng_mode=ol.RNG_GLOBAL, pick_mode=ol.PICK_CDF).run_cycle()
    assert o1.k_cycle == o2.k_cycle


def test_deck_error_messages():
    """the reference's messages where it has one (it prints and exits; the loader returns them)"""
    with pytest.raises(ValueError, match="Unknown source type"):
        mcb.Deck(xml=decks.slab(samples=10).replace("<point ", "<ring_z "))
    with pytest.raises(ValueError, match="Unknown nuclide"):
        mcb.Deck(xml=decks.slab(samples=10).replace('<nuclide name="nuc3" density="0.1"/>', '<nuclide name="nope" density="0.1"/>'))
    with pytest.raises(ValueError, match="ksearch and tdmc could not coexist"):   # setup.cpp:166-169
        mcb.Deck(xml=decks.gcr(samples=10).replace("</simulation>", '<tdmc time="1.0 2.0"/></simulation>'))
    with pytest.raises(ValueError, match="needs a time grid"):
        mcb.Deck(xml=decks.slab(samples=10).replace("</simulation>", '<tdmc/></simulation>'))
    with pytest.raises(ValueError, match="TRMM should be run in ksearch mode"):
        mcb.Deck(xml=decks.slab(samples=10) + '<trmm><cell name="slab 1"/><filter type="energy" grid="1 2"/></trmm>')
    assert mcb.Deck(xml=decks.gcr(samples=10, trmm=True), flags=mcb.IGNORE_TRMM).info["n_estimators"] == 0
    t = mcb.Deck(xml=decks.gcr(samples=10, trmm=True))
    assert t.info["trmm_present"] == 1 and t.info["n_estimators"] == 9 and t.info["n_tallies"] == 3500
    with pytest.raises(ValueError):
        mcb.Deck(io_dir="/nonexistent/dir")


@pytest.mark.skipif(not os.path.isdir("/root/reference/examples"), reason="reference tree not present")
@pytest.mark.parametrize("example", ["slab_analytic", "HEU_sphere_criticality", "shielding_vReduction", "UCube", "infinite_GCR_TRMM",
                                     "infinite_GCR_TRMM_100", "infinite_GCR_TRMM_critical", "infinite_GCR_TRMM_critical2",
                                     "infinite_GCR_Ttmp", "HEU_sphere_leakage", "infinite_GCR_TD", "infinite_GCR_TD_sub",
                                     "sphere_detection"])
def test_reference_example_decks_load_unchanged(example):
    """the reference's own input.xml files parse as they are, TRMM tally sets, time filters, time-dependent mode and
    particle comb included - every example of the tree, sphere_detection (whose <disk_z> source the reference itself
    rejects, setup.cpp:1051-1063) among them"""
    deck = mcb.Deck(io_dir="/root/reference/examples/" + example)
    i = deck.info
    assert i["n_sample"] > 0 and i["n_cells"] > 0 and i["n_sources"] > 0
    assert deck.mode == ("k-eigenvalue" if i["ksearch"] else "time-dependent" if "_TD" in example else "fixed source")


def test_estimator_layout_matches_reference_order():
    """[score][filter1][filter2] row-major (Estimator.cpp:280-295)"""
    deck = mcb.Deck(xml=decks.fixed_source_fissile(samples=10))
    est = {e["name"]: e for e in deck.estimators()}
    tl = est["core_tl"]
    assert tl["scores"] == ["flux", "fission", "nu-fission", "capture", "scatter", "total", "absorption"]
    assert [f["size"] for f in tl["filters"]] == [2, 6]
    assert tl["n_tallies"] == 7 * 2 * 6
    assert est["interface"]["scores"] == ["cross", "flux"] and est["interface"]["n_tallies"] == 4
    assert sum(e["n_tallies"] for e in est.values()) == deck.info["n_tallies"]


def _reference_bisect(grid, E):
    """binary_search of the reference (Algorithm.cpp:46-64), probe by probe, vectorised over E; valid for grids that
    are not ascending too (xs_library/005011.txt has two descents)"""
    left = np.zeros(E.shape, dtype=np.int64)
    right = np.full(E.shape, len(grid) - 1, dtype=np.int64)
    while True:
        live = left <= right
        if not live.any():
            return right
        mid = (left + right) // 2
        less = np.zeros(E.shape, dtype=bool)
        less[live] = grid[mid[live]] < E[live]
        left = np.where(live & less, mid + 1, left)
        right = np.where(live & ~less, mid - 1, right)


@pytest.mark.parametrize("name", sorted(gc.function_decks()))
def test_union_grid_indices_equal_reference_bisection(name):
    """the device lookup structure (union grid + map + hash, mcb_tables) yields, for every nuclide of every material,
    the row index the reference's per-nuclide bisection yields: every grid point and its +-1 ulp neighbours,
    duplicates, out-of-range energies, and the non-ascending B-11 grid"""
    deck = mcb.Deck(xml=gc.function_decks()[name])
    L = mcb.host_lib()
    L.mcbh_union_indices.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p]
    for m, zaids in enumerate(gc._MAT_ZAIDS[name]):
        grids = [gc.grid_energies(z) for z in zaids]
        allg = np.concatenate(grids)
        E = np.concatenate([gc.energies(name, m), allg, np.nextafter(allg, 0.0), np.nextafter(allg, np.inf),
                            np.array([0.0, -1.0, 1e-300, 1e300])])
        idx = np.zeros((E.size, len(zaids)), dtype=np.int32)
        stats = np.zeros(4, dtype=np.int64)
        nn = L.mcbh_union_indices(deck._h, m, E.ctypes.data, E.size, idx.ctypes.data, stats.ctypes.data)
        assert nn == len(zaids)
        for k, g in enumerate(grids):
            want = _reference_bisect(g, E)
            assert np.array_equal(idx[:, k], want), "material %d nuclide %s of %s" % (m, zaids[k], name)
        assert stats[0] <= allg.size and stats[3] >= 1


# ---- N > 1 data flow over gloo ----------------------------------------------------------------------------------
def _gloo_worker(rank, world, port, xml, cycles, q):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    deck = mcb.Deck(xml=xml)
    n = deck.info["n_sample"]
    b, c = mcb.shard_range(n, rank, world)
    orc = ol.Oracle(deck, rng_mode=ol.RNG_HISTORY, pick_mode=ol.PICK_FLOOR)
    orc.set_shard(b, c)
    out = []
    for _ in range(cycles):
        orc.transport_cycle()
        sums, counts = orc.partials()
        ts, tq = orc.tally_partials()
        sites, cells = orc.bank()
        # all-reduce of the sums (the library does this in exact integer arithmetic; doubles here)
        t = torch.from_numpy(np.concatenate([sums, counts.astype(np.float64), ts, tq]))
        dist.all_reduce(t)
        t = t.numpy()
        nt = ts.size
        # all-gather of the bank: counts first, then the padded slices, concatenated in rank order
        cnt = [torch.zeros(1, dtype=torch.int64) for _ in range(world)]
        dist.all_gather(cnt, torch.tensor([sites.shape[0]], dtype=torch.int64))
        m = max(int(x) for x in cnt)
        pad = torch.zeros((m, 9), dtype=torch.float64)
        pad[:sites.shape[0], :8] = torch.from_numpy(sites)
        pad[:sites.shape[0], 8] = torch.from_numpy(cells.astype(np.float64))
        parts = [torch.zeros((m, 9), dtype=torch.float64) for _ in range(world)]
        dist.all_gather(parts, pad)
        g = torch.cat([p[:int(k)] for p, k in zip(parts, cnt)]).numpy()
        r = orc.close_cycle(t[:5], t[5:9].astype(np.uint64), t[9:9 + nt], t[9 + nt:])
        orc.set_source_bank(np.ascontiguousarray(g[:, :8]), g[:, 8].astype(np.int32))
        out.append((r.k_cycle, int(r.n_sites), int(r.n_tracks), float(np.sum(g[:, 6]))))
    q.put((rank, out))
    dist.barrier()
    dist.destroy_process_group()


def test_gloo_world2_sharded_generations_match_one_rank():
    """two ranks, each transporting its shard of every generation, combined the way the library combines GPUs
    (SURVEY §8e): same fission bank (rank-ordered concatenation = canonical order) and same k as one rank"""
    import torch.multiprocessing as mp
    xml = decks.heu_sphere(samples=3001, active=2, passive=1)
    cycles = 3
    deck = mcb.Deck(xml=xml)
    one = ol.Oracle(deck, rng_mode=ol.RNG_HISTORY, pick_mode=ol.PICK_FLOOR)
    want = []
    for _ in range(cycles):
        r, sites, cells = one.run_cycle_keep_bank()
        want.append((r.k_cycle, int(r.n_sites), int(r.n_tracks), float(np.sum(sites[:, 6]))))
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 400
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, xml, cycles, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = dict(q.get(timeout=300) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert got[0] == got[1]                       # the close-out is identical on every rank
    for (k, ns, nt, es), (k1, ns1, nt1, es1) in zip(got[0], want):
        assert ns == ns1 and nt == nt1             # integer results do not depend on the number of ranks
        assert es == pytest.approx(es1, rel=1e-13)  # same sites, same order (summation order is the same too)
        assert k == pytest.approx(k1, rel=1e-12)   # double sums of two partials vs one running sum


def test_xml_reader_covers_the_grammar_pugixml_accepts():
    """the host keeps its own small XML reader (host/xml_lite.cpp) where the reference links pugixml: everything a deck
    author may legally write around the deck grammar must load the same problem - comments (also with markup inside),
    single-quoted attributes, the five named entities and character references, a DOCTYPE, CDATA, free white space
    inside tags, explicit end tags, a byte-order mark - and malformed input is an error with a line number, never a
    half-read deck"""
    base = decks.slab(samples=10)
    ref = mcb.Deck(xml=base).info
    same = {
        "comments": base.replace("<surfaces>", "<!-- a comment with <tags> & ampersands -->\n<surfaces><!-- inside -->"),
        "single quotes": base.replace('name="px1" x="0.0"', "name='px1' x='0.0'"),
        "entities": base.replace('name="Simple Slabs"', 'name="Simple &amp; &lt;Slabs&gt; &quot;&apos; &#65;&#x42;"'),
        "doctype": base.replace("<simulation>", "<!DOCTYPE simulation>\n<simulation>", 1),
        "cdata": base.replace("</simulation>", "<![CDATA[ junk <x> ]]></simulation>"),
        "white space": base.replace('<plane_x name="px1" x="0.0"/>', '<plane_x   name = "px1"\n   x = "0.0"   />'),
        "end tag": base.replace('<plane_x name="px1" x="0.0"/>', '<plane_x name="px1" x="0.0"></plane_x>'),
        "byte-order mark": "﻿" + base,
    }
    for what, xml in same.items():
        assert xml != base, what
        assert mcb.Deck(xml=xml).info == ref, what
    for what, xml, msg in [("missing end tag", base.replace("</cells>", ""), r"missing </cells> \(line \d+\)"),
                           ("mismatched end tag", base.replace("</cells>", "</cell>"), r"mismatched </cell>, open <cells> \(line \d+\)"),
                           ("unquoted attribute", base.replace('x="0.0"/>', "x=0.0/>", 1), r"value not quoted \(line \d+\)")]:
        with pytest.raises(ValueError, match=msg):
            mcb.Deck(xml=xml)


def _cross_neighbors(deck):
    """{(surface name, side): cell name or None} from mcbh_cross_neighbors (the table the walk kernel's crossing uses)"""
    L = mcb.host_lib()
    L.mcbh_cross_neighbors.argtypes = [C.c_void_p, C.c_void_p, C.c_int32]
    n = deck.info["n_surfaces"]
    out = np.full(2 * max(n, 1), -7, dtype=np.int32)
    assert L.mcbh_cross_neighbors(deck._h, out.ctypes.data, out.size) == 2 * max(n, 1)
    return {(deck.name(2, s), side): (deck.name(3, int(out[2 * s + k])) if out[2 * s + k] >= 0 else None)
            for s in range(n) for k, side in enumerate((-1, +1))}


def test_crossing_shortcut_is_taken_only_where_search_cell_has_one_possible_answer():
    """search_cell returns the FIRST cell in deck order that contains the point (general.cpp:26-34).  The kernel may
    take the cell behind a surface from a table only where that answer is forced: every cell before it holds the
    surface with the other sense, and the cell itself consists of that one surface"""
    # bare sphere + graveyard: leaving the sphere can only end in the graveyard; entering it is never tabulated
    # (the sphere cell comes first and consists of (sphere, -1) alone: also forced)
    t = _cross_neighbors(mcb.Deck(xml=decks.heu_sphere(samples=10)))
    (sname, _), = {k for k in t if k[1] == +1}
    assert t[(sname, +1)] is not None and t[(sname, -1)] is not None and t[(sname, +1)] != t[(sname, -1)]
    # slab: behind px1 (x < 0) lies "left outside", but slab2 precedes it in the deck and does not mention px1 at all, so
    # the side does not rule it out: not forced (the kernel searches, as before)
    t = _cross_neighbors(mcb.Deck(xml=decks.slab(samples=10)))
    assert t[("px1", -1)] is None and t[("px3", +1)] is None
    assert all(v is None for v in t.values())
    # the overlap deck: "catch" (one surface: py, -1) precedes the outside cells and mentions neither px1 nor px3:
    # nothing about the x planes is forced; y > 1000 rules "catch" out, but slab1 / slab2 do not mention py: not forced
    t = _cross_neighbors(mcb.Deck(xml=decks.slab_overlap(samples=10)))
    assert all(v is None for v in t.values())
    # a one-cell world behind a single plane, listed first: forced on its side only
    xml = decks.slab(samples=10).replace('<cells>', '<cells>\n    <cell name="beyond" importance="0.0">\n        <surface name="px3" sense="+1"/>\n    </cell>') \
        .replace('<cell name="right outside" importance="0.0">\n        <surface name="px3" sense="+1"/>\n    </cell>', "")
    t = _cross_neighbors(mcb.Deck(xml=xml))
    assert t[("px3", +1)] == "beyond" and t[("px3", -1)] is None and t[("px1", -1)] is None


def _surface_sides(xml, pts):
    """sign of Surface*::eval (Geometry.cpp:29-69) of every surface of the deck at every point: {name: array of -1/0/+1}"""
    import xml.etree.ElementTree as ET
    root = ET.fromstring("<r>" + xml.split("?>", 1)[1] + "</r>")
    x, y, z = pts.T
    out = {}
    for s in root.find("surfaces"):
        a = {k: float(v) for k, v in s.attrib.items() if k not in ("name", "bc")}
        e = {"plane_x": lambda: x - a["x"], "plane_y": lambda: y - a["y"], "plane_z": lambda: z - a["z"],
             "plane": lambda: a["a"] * x + a["b"] * y + a["c"] * z - a["d"],
             "sphere": lambda: (x - a["x"]) ** 2 + (y - a["y"]) ** 2 + (z - a["z"]) ** 2 - a["r"] ** 2,
             "cylinder_x": lambda: (y - a["y"]) ** 2 + (z - a["z"]) ** 2 - a["r"] ** 2,
             "cylinder_z": lambda: (x - a["x"]) ** 2 + (y - a["y"]) ** 2 - a["r"] ** 2}[s.tag]()
        out[s.attrib["name"]] = np.sign(e).astype(int)
    return out


@pytest.mark.parametrize("name", ["heu", "slab", "slab_overlap", "shield", "fsf", "ucube", "sphere_det", "leak"])
def test_crossing_shortcut_agrees_with_search_cell_everywhere(name):
    """property behind the walk kernel's crossing shortcut: wherever the table names a cell for (surface, side), search_cell
    (mcbh_search_cell: general.cpp:26-34 on the host) returns that very cell at EVERY point strictly on that side - checked on
    2e4 random points per deck, half of them close to the surfaces' scale"""
    xml = {"heu": decks.heu_sphere(samples=10), "slab": decks.slab(samples=10), "slab_overlap": decks.slab_overlap(samples=10),
           "shield": decks.shielding(samples=10), "fsf": decks.fixed_source_fissile(samples=10), "ucube": decks.ucube(samples=10),
           "sphere_det": decks.sphere_detection(samples=10), "leak": decks.heu_leakage(samples=10)}[name]
    deck = mcb.Deck(xml=xml)
    table = _cross_neighbors(deck)
    rng = np.random.default_rng(7)
    pts = np.concatenate([rng.uniform(-30, 30, (10000, 3)), rng.uniform(-1200, 1200, (2000, 3)), rng.normal(0, 6, (8000, 3))])
    sides = _surface_sides(xml, pts)
    L = mcb.host_lib()
    L.mcbh_search_cell.argtypes = [C.c_void_p, C.c_double, C.c_double, C.c_double]
    found = np.array([L.mcbh_search_cell(deck._h, *p) for p in pts])
    cells = {deck.name(3, i): i for i in range(deck.info["n_cells"])}
    n_checked = 0
    for (sname, side), cname in table.items():
        if cname is None:
            continue
        on_side = sides[sname] == side
        n_checked += int(on_side.sum())
        assert np.all(found[on_side] == cells[cname]), (sname, side, cname)
    if name in ("heu", "leak"):
        assert n_checked > 10000   # both sides of the sphere are forced there
