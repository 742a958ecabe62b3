"""Output contract: the host program's own HDF5 writer (mc_old_b200/host/h5lite.cpp) against
(i) the dataset tree the compiled reference wrote for the same decks (tests/golden/runs.json, recorded through the
H5Cpp stand-in by tests/golden/make_golden.py) and (ii) real HDF5 files: the reference's committed examples/*/output.h5
are read with the same minimal reader (tests/h5mini.py) that reads ours."""
import ctypes as C
import glob
import json
import os
import re
import subprocess

import numpy as np
import pytest

import golden_cases as gc
import h5mini
import mc_old_b200 as mcb
from mc_old_b200 import decks

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = json.load(open(os.path.join(ROOT, "tests", "golden", "runs.json")))


def _write(deck, path, n_track=12345):
    """output.h5 of `deck` with recognisable synthetic results; returns what was written"""
    L = mcb.host_lib()
    L.mcbh_write_output.argtypes = [C.c_void_p, C.c_char_p, C.c_uint64, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p,
                                    C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_int64]
    i = deck.info
    nc, na, nt = i["n_cycle"], i["n_cycle"] - i["n_passive"], i["n_tallies"]
    kc = 1.0 + 0.001 * np.arange(nc); hc = 0.25 + 0.01 * np.arange(nc)
    ka = 0.9 + 0.002 * np.arange(na); ku = 0.01 / (1.0 + np.arange(na))
    tm = 100.0 + np.arange(nt, dtype=np.float64); tu = 0.5 + np.arange(nt, dtype=np.float64)
    rc = L.mcbh_write_output(deck._h, path.encode(), n_track, kc.ctypes.data, hc.ctypes.data, nc, ka.ctypes.data,
                             ku.ctypes.data, na, tm.ctypes.data, tu.ctypes.data, nt)
    assert rc == 0, L.mcbh_last_error().decode()
    return dict(k_cycle=kc, H_cycle=hc, k_avg=ka, k_uncer=ku, mean=tm, uncer=tu)


@pytest.mark.parametrize("name", sorted(GOLDEN))
def test_output_tree_matches_the_reference_run(name, tmp_path):
    """same dataset paths and element counts as the reference wrote for the deck"""
    xml, _patched = gc.run_decks()[name]
    deck = mcb.Deck(xml=xml)
    path = str(tmp_path / "output.h5")
    w = _write(deck, path)
    f = h5mini.File(path)
    ours = {p: n for p, n in f.root.walk() if n.kind == "dataset"}
    want = {k: v for k, v in GOLDEN[name].items() if k.startswith("/")}
    assert sorted(ours) == sorted(want)
    for p, v in want.items():
        n = ours[p]
        count = int(np.prod(n.shape)) if n.shape else 1
        assert count == (len(v) if isinstance(v, list) else 1), p
        if p.startswith("/summary/") and p.split("/")[-1] in ("Ncycle", "Nsample", "Npassive"):
            assert n.dtype == "uint" and n.value == v[0]
        elif p == "/summary/mode":
            assert n.dtype == "vstr" and n.value == v
        elif p != "/summary/Ntrack":
            assert n.dtype == "f64"
    assert f.root["summary/Ntrack"].value == 12345
    assert f.root["summary/survival_roulette/wr"].value == float.fromhex(GOLDEN[name]["/summary/survival_roulette/wr"][0])
    if deck.info["ksearch"]:
        assert np.array_equal(f.root["ksearch/k_cycle"].value, w["k_cycle"])
        assert np.array_equal(f.root["ksearch/H_cycle"].value, w["H_cycle"])
        assert np.array_equal(f.root["ksearch/k_active/mean"].value, w["k_avg"])
        assert f.root["ksearch/mean"].value == w["k_avg"][-1] and f.root["ksearch/uncertainty"].value == w["k_uncer"][-1]


def test_trm_assembly_matches_the_reference(tmp_path):
    """report.cpp:53-157: TRM, inverse_speed, C_initial, psi_initial computed from the reference's own tally means
    equal what the reference wrote, bit for bit"""
    import report_order
    rec = {k: (np.array([float.fromhex(x) for x in v]) if k.startswith("/") and isinstance(v, list) and v and isinstance(v[0], str) else v)
           for k, v in GOLDEN["gcr_trmm"].items()}
    deck = mcb.Deck(xml=gc.run_decks()["gcr_trmm"][0])
    mean, _ = report_order.flatten(deck, rec)
    L = mcb.host_lib()
    L.mcbh_trm_assemble.argtypes = [C.c_void_p] + [C.c_void_p] * 5
    G = 20
    TRM = np.zeros((G + 6) ** 2); inv = np.zeros(G); Ci = np.zeros(6); psi = np.zeros(G)
    assert L.mcbh_trm_assemble(deck._h, mean.ctypes.data, TRM.ctypes.data, inv.ctypes.data, Ci.ctypes.data, psi.ctypes.data) == G
    for got, key in ((TRM, "/TRM"), (inv, "/inverse_speed"), (Ci, "/C_initial"), (psi, "/psi_initial")):
        want = rec[key]
        assert np.array_equal(np.isnan(got), np.isnan(want)), key
        assert np.array_equal(got[~np.isnan(got)], want[~np.isnan(want)]), key
    heu = mcb.Deck(xml=decks.heu_sphere(samples=10))
    assert L.mcbh_trm_assemble(heu._h, mean.ctypes.data, TRM.ctypes.data, inv.ctypes.data, Ci.ctypes.data, psi.ctypes.data) == -1


def test_estimator_groups_layout_and_attributes(tmp_path):
    """/<estimator> attr indexing, /<estimator>/<filter> grids with attr unit, /<estimator>/<score>/{mean,uncertainty}
    shaped by the filter sizes, values in the flat [score][filters...] order (Estimator.cpp:368-422)"""
    deck = mcb.Deck(xml=decks.fixed_source_fissile(samples=10))
    path = str(tmp_path / "output.h5")
    w = _write(deck, path)
    f = h5mini.File(path)
    assert f.root["summary/mode"].value == "fixed source" and "ksearch" not in f.root.children
    for e in deck.estimators():
        g = f.root[e["name"]]
        assert g.kind == "group"
        names = {0: "surface", 1: "cell", 2: "energy", 3: "energy_initial", 4: "time"}
        units = {0: "id#", 1: "id#", 2: "eV", 3: "eV", 4: "s"}
        assert g.attrs["indexing"] == "".join("[%s]" % names[fl["type"]] for fl in e["filters"])
        shape = tuple(fl["size"] for fl in e["filters"])
        for fl in e["filters"]:
            d = g[names[fl["type"]]]
            assert d.attrs["unit"] == units[fl["type"]] and d.shape == (len(fl["grid"]),)
            assert np.array_equal(d.value, np.array(fl["grid"]))
        per = int(np.prod(shape))
        for k, s in enumerate(e["scores"]):
            t0 = e["tally_begin"] + k * per
            assert g[s + "/mean"].shape == shape
            assert np.array_equal(g[s + "/mean"].value.ravel(), w["mean"][t0:t0 + per])
            assert np.array_equal(g[s + "/uncertainty"].value.ravel(), w["uncer"][t0:t0 + per])


def test_many_children_and_odd_names(tmp_path):
    """more than 8 entries in a group (several symbol nodes under one B-tree node), names with spaces and '#'"""
    xml = decks.fixed_source_fissile(samples=10)
    est = "".join('<estimator name="Estimator #%02d x" scores="flux"><cell name="core"/></estimator>' % i for i in range(23))
    xml = xml.replace("</estimators>", est + "</estimators>")
    deck = mcb.Deck(xml=xml)
    path = str(tmp_path / "output.h5")
    _write(deck, path)
    f = h5mini.File(path)
    for i in range(23):
        assert f.root["Estimator #%02d x/flux/mean" % i].shape == (1,)
    assert len(f.root.children) == 23 + len(mcb.Deck(xml=decks.fixed_source_fissile(samples=10)).estimators()) + 1


@pytest.mark.skipif(not os.path.isdir("/root/reference/examples"), reason="reference tree not present")
def test_reader_reads_files_written_by_the_real_hdf5_library():
    """validates the reader (and so the format knowledge the writer rests on) on the reference's committed outputs"""
    files = sorted(glob.glob("/root/reference/examples/*/output.h5"))
    assert files
    for path in files:
        f = h5mini.File(path)
        s = f.root["summary"]
        assert s["mode"].dtype == "vstr" and s["mode"].value in ("fixed source", "k-eigenvalue", "time-dependent")
        assert s["Nsample"].dtype == "uint" and s["Nsample"].value > 0
    f = h5mini.File("/root/reference/examples/infinite_GCR_TRMM/output.h5")
    assert f.root["TRM_simple"].attrs["indexing"] == "[cell][energy]"
    assert f.root["TRM_simple/flux/mean"].shape == (1, 6)
    assert f.root["ksearch/k_cycle"].shape == (110,)


@pytest.mark.skipif(not os.path.isdir("/root/reference/examples"), reason="reference tree not present")
def test_same_structures_as_the_real_library_writes():
    """byte-level facts shared with a libhdf5-written file: superblock v0 fields, root entry cache type, v1 object
    header of a group (one symbol-table message), TREE node size"""
    ref = open("/root/reference/examples/infinite_GCR_TRMM/output.h5", "rb").read()
    deck = mcb.Deck(xml=decks.heu_sphere(samples=10, active=1, passive=1))
    import tempfile
    with tempfile.TemporaryDirectory() as d:
        _write(deck, os.path.join(d, "o.h5"))
        ours = open(os.path.join(d, "o.h5"), "rb").read()
    assert ours[:24] == ref[:24]                      # signature, versions, sizes, K values, consistency flags
    assert ours[24:40] == ref[24:40]                  # base address, free-space address
    assert ours[48:56] == ref[48:56]                  # driver info address
    assert ours[72:76] == ref[72:76]                  # root entry: cache type 1
    import struct
    eof = struct.unpack_from("<Q", ours, 40)[0]
    assert eof == len(ours)
    ro = struct.unpack_from("<Q", ours, 64)[0]
    rr = struct.unpack_from("<Q", ref, 64)[0]
    assert ours[ro:ro + 16] == ref[rr:rr + 16]        # object header prefix of the root group
    assert ours[ro + 16:ro + 24] == ref[rr + 16:rr + 24]  # symbol-table message header


@pytest.mark.gpu
def test_host_program_cli_and_output(tmp_path):
    """MCB.exe <dir>: the reference's banners and cycle lines (Main.cpp:22-28, Estimator.cpp:536-554), output.h5
    equal to what the C-ABI returns for the same deck"""
    exe = os.path.join(ROOT, "mc_old_b200", "MCB.exe")
    assert os.path.exists(exe), "MCB.exe is not built (run __graft_entry__.build())"
    xml = decks.heu_sphere(samples=20000, active=3, passive=2, entropy=True, estimators=True)
    d = str(tmp_path)
    decks.write(d, xml)
    env = dict(os.environ, MCB_XS_LIBRARY=mcb.default_xs_dir())
    out = subprocess.run([exe, d], capture_output=True, text=True, env=env, timeout=300)
    assert out.returncode == 0, out.stdout + out.stderr
    assert out.stdout.startswith("\nSimulation setup done,\nNow running the simulation...\n\n")
    assert out.stdout.endswith("Simulation done!\n\nReporting simulation output...\nSimulation output done!\n")
    lines = [l for l in out.stdout.split("\n") if re.match(r"^\d+   ", l)]
    assert len(lines) == 5
    assert re.match(r"^1   [0-9.e+-]+   \([0-9.e+-]+\)$", lines[0])
    assert re.match(r"^5   [0-9.e+-]+   [0-9.e+-]+   \+/-   [0-9.e+-]+   \([0-9.e+-]+\)$", lines[4])
    f = h5mini.File(os.path.join(d, "output.h5"))
    deck = mcb.Deck(xml=xml)
    ctx = mcb.Context(deck, device=0)
    rs = [ctx.run_cycle() for _ in range(5)]
    mean, uncer = ctx.tallies()
    ctx.close()
    assert np.array_equal(f.root["ksearch/k_cycle"].value, np.array([r.k_cycle for r in rs]))
    assert np.array_equal(f.root["ksearch/H_cycle"].value, np.array([r.H for r in rs]))
    assert f.root["ksearch/mean"].value == rs[-1].k_avg and f.root["ksearch/uncertainty"].value == rs[-1].k_uncer
    assert f.root["summary/Ntrack"].value == sum(r.n_tracks for r in rs)
    # tallies: the cycle sums are accumulated with floating-point reductions in no fixed order (rounding-level spread)
    assert np.allclose(f.root["sphere_rates/flux/mean"].value.ravel(), mean[0:7], rtol=1e-12, atol=0)
    assert np.allclose(f.root["leak/cross/uncertainty"].value.ravel(), uncer[30:31], rtol=1e-9, atol=0)
    # no argument: the reference's message and a failure exit code (Main.cpp:11-14)
    out = subprocess.run([exe], capture_output=True, text=True)
    assert out.returncode != 0 and "[ERROR] Please provide input.xml directory..." in out.stdout


@pytest.mark.gpu
def test_host_program_on_two_gpus(tmp_path):
    """MCB.exe started once per GPU (RANK / WORLD_SIZE / LOCAL_RANK, NCCL id through a file): same cycle lines and
    the same output.h5 as the single-GPU run, bit for bit (k, H, Ntrack; tallies to rounding)"""
    if mcb.cuda_lib().mcb_device_count() < 2:
        pytest.skip("needs 2 GPUs")
    exe = os.path.join(ROOT, "mc_old_b200", "MCB.exe")
    xml = decks.heu_sphere(samples=40000, active=2, passive=2, entropy=True, estimators=True)
    d1, d2 = str(tmp_path / "one"), str(tmp_path / "two")
    decks.write(d1, xml); decks.write(d2, xml)
    env = dict(os.environ, MCB_XS_LIBRARY=mcb.default_xs_dir())
    one = subprocess.run([exe, d1], capture_output=True, text=True, env=env, timeout=300)
    assert one.returncode == 0, one.stdout + one.stderr
    idf = str(tmp_path / "nccl_id")
    procs = [subprocess.Popen([exe, d2], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True,
                              env=dict(env, RANK=str(r), WORLD_SIZE="2", LOCAL_RANK=str(r), MCB_ID_FILE=idf)) for r in range(2)]
    outs = [p.communicate(timeout=300)[0] for p in procs]
    assert all(p.returncode == 0 for p in procs), outs
    lines = [l for l in outs[0].splitlines() if not l.startswith("NCCL version")]  # libnccl's own banner
    assert lines == one.stdout.splitlines(), (outs[0], one.stdout)
    assert outs[1] == ""
    a, b = h5mini.File(os.path.join(d1, "output.h5")), h5mini.File(os.path.join(d2, "output.h5"))
    for key in ("ksearch/k_cycle", "ksearch/H_cycle", "ksearch/k_active/mean", "ksearch/k_active/uncertainty"):
        assert np.array_equal(a.root[key].value, b.root[key].value), key
    assert a.root["summary/Ntrack"].value == b.root["summary/Ntrack"].value
    assert np.allclose(a.root["sphere_rates/flux/mean"].value, b.root["sphere_rates/flux/mean"].value, rtol=1e-12, atol=0)


@pytest.mark.skipif(not os.path.isdir("/root/reference/examples"), reason="reference tree not present")
@pytest.mark.parametrize("example,newer", [("infinite_GCR_TRMM_critical", []), ("infinite_GCR_TRMM_critical2", []),
                                           ("infinite_GCR_TRMM_100", ["/C_initial", "/psi_initial"])])
def test_output_tree_equals_the_committed_reference_outputs(example, newer, tmp_path):
    """the reference's own examples/<deck>/output.h5 (written by the real HDF5 library in the author's runs) and the
    file our writer produces for the same input.xml hold the same objects: every path, dtype and shape (123-125
    objects: /summary, /ksearch, nine TRMM estimator groups with filters and scores, /TRM, /inverse_speed, ...).
    `newer` = datasets the current reference source writes but the committed file of that deck predates.
    (examples/infinite_GCR_TRMM/output.h5 is left out: it was produced with another deck, 6 groups instead of 20.)"""
    d = "/root/reference/examples/" + example
    deck = mcb.Deck(io_dir=d)
    path = str(tmp_path / "output.h5")
    _write(deck, path)
    ref, ours = h5mini.tree(os.path.join(d, "output.h5")), h5mini.tree(path)
    assert sorted(set(ref) - set(ours)) == []
    assert sorted(set(ours) - set(ref)) == newer
    assert all(ref[k] == ours[k] for k in ref)
    r, o = h5mini.File(os.path.join(d, "output.h5")), h5mini.File(path)
    for p, n in r.root.walk():
        assert o.root[p].attrs.keys() == n.attrs.keys(), p
        if n.kind == "group" and "indexing" in n.attrs:
            assert o.root[p].attrs["indexing"] == n.attrs["indexing"]
        if n.kind == "dataset" and "unit" in n.attrs:
            assert o.root[p].attrs["unit"] == n.attrs["unit"]
            assert np.allclose(o.root[p].value, n.value, rtol=1e-12), p   # the filter grids themselves
