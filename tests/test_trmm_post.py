"""The reference's second program, TRMM.exe (TRMM.cpp:10-81): eigen-pairs of the Transition Rate Matrix, forward and
adjoint, read from output.h5 and written to output_TRMM.h5.  Host-only (MCB_TRMM.exe / mcbh_trmm_postprocess).

Checked against (i) the reference's own committed result for its infinite_GCR_TRMM example (tests/golden/trmm_eigen.npz,
made by tests/golden/make_trmm_eigen_golden.py from files the reference's TRMM.exe — Eigen::EigenSolver — wrote),
(ii) LAPACK through numpy on matrices that stress the iteration, (iii) the defining property A v = alpha v."""
import ctypes as C
import json
import os
import shutil
import subprocess

import numpy as np
import pytest

import golden_cases as gc
import h5mini
import mc_old_b200 as mcb
import report_order

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = np.load(os.path.join(ROOT, "tests", "golden", "trmm_eigen.npz"))
EXE = os.path.join(ROOT, "mc_old_b200", "MCB_TRMM.exe")


def eigen(A):
    L = mcb.host_lib()
    L.mcbh_eigen_general.argtypes = [C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p]
    A = np.ascontiguousarray(A, dtype=np.float64)
    n = A.shape[0]
    w = np.zeros(2 * n); V = np.zeros(2 * n * n)
    rc = L.mcbh_eigen_general(n, A.ctypes.data, w.ctypes.data, V.ctypes.data)
    assert rc == 0, L.mcbh_last_error().decode()
    return w.view(np.complex128).copy(), V.view(np.complex128).reshape(n, n).copy()


def check_conventions(A, w, V):
    n = len(w)
    # sorted by descending real part, conjugate pairs adjacent and exactly conjugate, positive imaginary part first
    assert all(w[i].real >= w[i + 1].real for i in range(n - 1))
    i = 0
    while i < n:
        if w[i].imag != 0.0:
            assert w[i].imag > 0 and w[i + 1] == np.conj(w[i]) and np.array_equal(V[:, i + 1], np.conj(V[:, i]))
            i += 2
        else:
            assert np.all(V[:, i].imag == 0.0)
            i += 1
    assert np.allclose(np.linalg.norm(V, axis=0), 1.0, rtol=0, atol=1e-14)
    big = V[np.argmax(np.abs(V), axis=0), np.arange(n)]
    assert np.all(big.real > 0) and np.allclose(big.imag, 0.0, atol=1e-15)
    # the defining property, relative to the size of the matrix
    assert np.max(np.linalg.norm(A @ V - V * w[None, :], axis=0)) <= 1e-13 * max(np.linalg.norm(A, 2), 1e-300) * n


@pytest.mark.parametrize("n", [1, 2, 3, 5, 12, 26, 60, 106])
@pytest.mark.parametrize("kind", ["dense", "graded", "triangular", "reducible", "rotation_blocks"])
def test_eigen_solver_against_lapack(n, kind):
    rng = np.random.default_rng(1000 * n + len(kind))
    A = rng.standard_normal((n, n))
    if kind == "graded":            # rows spanning 9 decades, like a TRM (1e7 /s fast groups, 1e-2 /s precursors)
        A = A * np.exp(rng.uniform(-10, 10, (n, 1)))
    elif kind == "triangular":
        A = np.triu(A)
    elif kind == "reducible" and n > 2:
        A[n // 2:, :n // 2] = 0.0
    elif kind == "rotation_blocks":  # conjugate pairs only (and one real eigenvalue when n is odd)
        B = np.zeros((n, n))
        for k in range(0, n - 1, 2):
            t = rng.uniform(0.1, 3.0); s = rng.uniform(0.5, 2.0)
            B[k:k + 2, k:k + 2] = s * np.array([[np.cos(t), -np.sin(t)], [np.sin(t), np.cos(t)]])
        if n % 2:
            B[-1, -1] = 0.3
        Q, _ = np.linalg.qr(A)
        A = Q @ B @ Q.T
    w, V = eigen(A)
    check_conventions(A, w, V)
    ref = np.linalg.eigvals(A)
    # condition-aware tolerance: LAPACK's own error on the graded matrices is ~1e-9 relative
    tol = 1e-7 if kind == "graded" else 1e-9
    a, b = np.sort_complex(w), np.sort_complex(ref)
    assert np.max(np.abs(a - b) / np.maximum(np.abs(b), 1e-300)) < tol


def test_solver_rejects_what_it_cannot_solve():
    L = mcb.host_lib()
    L.mcbh_eigen_general.argtypes = [C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p]
    A = np.array([[1.0, np.nan], [0.0, 1.0]]); w = np.zeros(4); V = np.zeros(8)
    assert L.mcbh_eigen_general(2, A.ctypes.data, w.ctypes.data, V.ctypes.data) == -1
    assert b"non-finite" in L.mcbh_last_error()
    assert L.mcbh_eigen_general(0, A.ctypes.data, w.ctypes.data, V.ctypes.data) == -1


def same_direction(u, v):
    return abs(np.vdot(u, v)) / (np.linalg.norm(u) * np.linalg.norm(v))


def adjoint_matrix(TRM, inv):
    """TRMM.cpp:47-58"""
    G = len(inv)
    A = TRM.copy()
    A[:G, :] *= inv[:, None]
    A = A.T.copy()
    A[:G, :] /= inv[:, None]
    return A


def test_eigenpairs_of_the_reference_example():
    """same TRM in, the eigen-pairs the reference's TRMM.exe wrote out: eigenvalues to the accuracy of Eigen's unbalanced
    iteration (2.6e-9 against LAPACK on this matrix), eigenvectors equal up to the (arbitrary) complex scale"""
    TRM, inv = GOLD["TRM"], GOLD["inverse_speed"]
    for A, kw, kv in ((TRM, "alpha", "phi_mode"), (adjoint_matrix(TRM, inv), "alpha_adj", "phi_mode_adj")):
        w, V = eigen(A)
        check_conventions(A, w, V)
        w_ref, V_ref = GOLD[kw].ravel(), GOLD[kv]
        assert np.all(w.imag == 0.0) and np.all(w_ref.imag == 0.0)  # this TRM has a real spectrum
        order = np.argsort(-w_ref.real)
        assert np.max(np.abs(w - w_ref[order]) / np.abs(w_ref[order])) < 2e-8
        for j, jr in enumerate(order):
            assert same_direction(V[:, j], V_ref[:, jr]) > 1 - 1e-9, (kw, j)


def write_trmm_run(tmp_path):
    """output.h5 of the GCR TRMM deck (BASELINE configs[2]) written by the host library from the tally means of a
    reference-identical run at 4000 histories per generation (tests/golden/gcr_trmm_global_4000.npz; the 40-history
    golden run leaves energy groups unvisited and its TRM holds 0 / 0)"""
    from mc_old_b200 import decks
    g = np.load(os.path.join(ROOT, "tests", "golden", "gcr_trmm_global_4000.npz"))
    deck = mcb.Deck(xml=decks.gcr(samples=int(g["samples"]), active=int(g["active"]), passive=int(g["passive"]), trmm=True))
    mean, uncer = np.ascontiguousarray(g["mean"]), np.ascontiguousarray(g["uncer"])
    L = mcb.host_lib()
    L.mcbh_write_output.argtypes = [C.c_void_p, C.c_char_p, C.c_uint64, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p,
                                    C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_int64]
    i = deck.info
    nc, na = i["n_cycle"], i["n_cycle"] - i["n_passive"]
    kc = np.ones(nc); hc = np.ones(nc); ka = np.ones(na); ku = np.zeros(na)
    path = str(tmp_path / "output.h5")
    rc = L.mcbh_write_output(deck._h, path.encode(), 1, kc.ctypes.data, hc.ctypes.data, nc, ka.ctypes.data, ku.ctypes.data, na,
                             mean.ctypes.data, uncer.ctypes.data, len(mean))
    assert rc == 0, L.mcbh_last_error().decode()
    return path


def test_unvisited_energy_groups_are_an_error_not_garbage(tmp_path):
    """the 40-history golden run leaves groups unvisited: TRM holds 0 / 0 (the reference writes the NaNs and hands them to
    Eigen); the post-processor says so"""
    runs = json.load(open(os.path.join(ROOT, "tests", "golden", "runs.json")))
    rec = {k: (np.array([float.fromhex(x) for x in v]) if k.startswith("/") and isinstance(v, list) and v and isinstance(v[0], str) else v)
           for k, v in runs["gcr_trmm"].items()}
    deck = mcb.Deck(xml=gc.run_decks()["gcr_trmm"][0])
    mean, uncer = report_order.flatten(deck, rec)
    import test_output
    L = mcb.host_lib()
    i = deck.info
    nc, na = i["n_cycle"], i["n_cycle"] - i["n_passive"]
    kc = np.ones(nc); ka = np.ones(na)
    path = str(tmp_path / "output.h5")
    L.mcbh_write_output.argtypes = [C.c_void_p, C.c_char_p, C.c_uint64, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p,
                                    C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_int64]
    assert L.mcbh_write_output(deck._h, path.encode(), 1, kc.ctypes.data, kc.ctypes.data, nc, ka.ctypes.data, ka.ctypes.data, na,
                               mean.ctypes.data, uncer.ctypes.data, len(mean)) == 0
    r = subprocess.run([EXE, path], capture_output=True, text=True)
    assert r.returncode == 1 and "non-finite" in r.stderr


def test_postprocess_program_end_to_end(tmp_path):
    path = write_trmm_run(tmp_path)
    r = subprocess.run([EXE, path], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    run, out = h5mini.File(path), h5mini.File(str(tmp_path / "output_TRMM.h5"))
    TRM, inv = run.root["TRM"].value, run.root["inverse_speed"].value
    N = TRM.shape[0]
    assert N == len(inv) + 6 == 26
    assert sorted(out.root.children) == ["alpha", "alpha_adj", "phi_mode", "phi_mode_adj"]
    for k, shape in (("alpha", (N, 1)), ("alpha_adj", (N, 1)), ("phi_mode", (N, N)), ("phi_mode_adj", (N, N))):
        assert out.root[k].dtype == "c128" and out.root[k].shape == shape
    finite = np.isfinite(TRM).all()
    assert finite
    for A, kw, kv in ((TRM, "alpha", "phi_mode"), (adjoint_matrix(TRM, inv), "alpha_adj", "phi_mode_adj")):
        w, V = out.root[kw].value.ravel(), out.root[kv].value
        check_conventions(A, w, V)
        ref = np.linalg.eigvals(A)
        a, b = np.sort_complex(w), np.sort_complex(ref)
        assert np.max(np.abs(a - b) / np.abs(b)) < 1e-7
    # forward and adjoint spectra coincide (the adjoint matrix is similar to the transpose)
    a, b = np.sort_complex(out.root["alpha"].value.ravel()), np.sort_complex(out.root["alpha_adj"].value.ravel())
    assert np.max(np.abs(a - b) / np.abs(a)) < 1e-7
    # the reference's consumer (examples/infinite_GCR_TRMM/plot.py) pairs mode i of the two sorted spectra
    # and divides by gamma_i = sum_g phi_adj[g][i] * (1/v_g or 1) * phi[g][i]: bi-orthogonality keeps gamma away from 0
    alpha, phi, phi_adj = out.root["alpha"].value.ravel(), out.root["phi_mode"].value, out.root["phi_mode_adj"].value
    metric = np.concatenate([inv, np.ones(6)])
    gram = phi_adj.T @ (metric[:, None] * phi)
    off = gram - np.diag(np.diag(gram))
    assert np.max(np.abs(off)) < 1e-6 * np.min(np.abs(np.diag(gram))) or np.max(np.abs(off) / np.sqrt(np.abs(np.outer(np.diag(gram), np.diag(gram))))) < 1e-5


def test_postprocess_errors(tmp_path):
    r = subprocess.run([EXE], capture_output=True, text=True)
    assert r.returncode == 2 and "usage" in r.stderr
    r = subprocess.run([EXE, str(tmp_path / "nothing.h5")], capture_output=True, text=True)
    assert r.returncode == 1 and "[ERROR]" in r.stderr and "cannot open" in r.stderr
    (tmp_path / "junk.h5").write_bytes(b"not hdf5" * 100)
    r = subprocess.run([EXE, str(tmp_path / "junk.h5")], capture_output=True, text=True)
    assert r.returncode == 1 and "not an HDF5 file" in r.stderr
    # a run without a TRMM tally set has no TRM
    from mc_old_b200 import decks
    import test_output
    deck = mcb.Deck(xml=decks.heu_sphere(samples=10))
    test_output._write(deck, str(tmp_path / "output.h5"))
    r = subprocess.run([EXE, str(tmp_path / "output.h5")], capture_output=True, text=True)
    assert r.returncode == 1 and 'no dataset "TRM"' in r.stderr
    assert not (tmp_path / "output_TRMM.h5").exists()


@pytest.mark.skipif(not os.path.isdir("/root/reference/examples/infinite_GCR_TRMM"), reason="reference tree not present")
def test_reads_the_output_file_the_real_hdf5_library_wrote(tmp_path):
    """the reader behind MCB_TRMM.exe on a libhdf5-written output.h5 (the reference's committed example), result against
    the output_TRMM.h5 committed next to it"""
    shutil.copy("/root/reference/examples/infinite_GCR_TRMM/output.h5", tmp_path / "output.h5")
    L = mcb.host_lib()
    L.mcbh_trmm_postprocess.argtypes = [C.c_char_p]
    assert L.mcbh_trmm_postprocess(str(tmp_path / "output.h5").encode()) == 0, L.mcbh_last_error().decode()
    ours = h5mini.File(str(tmp_path / "output_TRMM.h5"))
    theirs = h5mini.File("/root/reference/examples/infinite_GCR_TRMM/output_TRMM.h5")
    for k in ("alpha", "alpha_adj"):
        a, b = np.sort_complex(ours.root[k].value.ravel()), np.sort_complex(theirs.root[k].value.ravel())
        assert ours.root[k].shape == theirs.root[k].shape
        assert np.max(np.abs(a - b) / np.abs(b)) < 2e-8
    for k in ("phi_mode", "phi_mode_adj"):
        assert ours.root[k].shape == theirs.root[k].shape and ours.root[k].dtype == theirs.root[k].dtype == "c128"


@pytest.mark.gpu
def test_both_programs_in_sequence_on_the_gpu(tmp_path):
    """the reference's workflow for BASELINE configs[2]: MCB.exe <dir> (the TRMM run, on the GPU), then MCB_TRMM.exe
    <dir>/output.h5.  The spectrum is checked against the one of the reference-identical 4000-history run: the six
    precursor modes sit at minus the decay constants whatever the statistics, the prompt modes move with the noise of
    the tallies"""
    from mc_old_b200 import decks
    d = str(tmp_path)
    decks.write(d, decks.gcr(samples=20000, active=4, passive=2, trmm=True))
    env = dict(os.environ, MCB_XS_LIBRARY=mcb.default_xs_dir())
    out = subprocess.run([os.path.join(ROOT, "mc_old_b200", "MCB.exe"), d], capture_output=True, text=True, env=env, timeout=600)
    assert out.returncode == 0, out.stdout + out.stderr
    r = subprocess.run([EXE, os.path.join(d, "output.h5")], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    run, post = h5mini.File(os.path.join(d, "output.h5")), h5mini.File(os.path.join(d, "output_TRMM.h5"))
    TRM = run.root["TRM"].value
    w, V = post.root["alpha"].value.ravel(), post.root["phi_mode"].value
    check_conventions(TRM, w, V)
    gold_dir = tmp_path / "gold"
    gold_dir.mkdir()
    gpath = write_trmm_run(gold_dir)
    L = mcb.host_lib()
    L.mcbh_trmm_postprocess.argtypes = [C.c_char_p]
    assert L.mcbh_trmm_postprocess(gpath.encode()) == 0
    wg = h5mini.File(str(gold_dir / "output_TRMM.h5")).root["alpha"].value.ravel()
    assert len(w) == len(wg) == 26
    # both sorted by descending real part: the slowest seven modes (six precursor groups and the fundamental prompt
    # mode of this subcritical medium) are well separated and statistically stiff
    assert np.allclose(w[:6].real, wg[:6].real, rtol=0.05) and np.all(w[:7].imag == 0.0)
    assert abs(w[6].real / wg[6].real - 1.0) < 0.25  # alpha_0 ~ (k - 1) / Lambda: 1 - k = 0.12 is known to ~4 % in either run
    assert np.all(w.real < 0.0)  # k = 0.88: every mode decays
