"""ctypes bindings for the CHECKERS (test infrastructure only):

* oracle/libmc_oracle.so   — our C restatement of the reference algorithm (oracle/mc_oracle.c)
* oracle/_ref/libref_harness.so, oracle/_ref/MC_ref[_patched] — the reference itself, compiled from
  /root/reference by oracle/build_ref.py (present in the container and, prebuilt, on the GPU box)

Nothing under mc_old_b200/ imports this module.
"""
import ctypes as C
import os
import re
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_LIB = os.path.join(ROOT, "oracle", "libmc_oracle.so")
REF_DIR = os.path.join(ROOT, "oracle", "_ref")
REF_HARNESS = os.path.join(REF_DIR, "libref_harness.so")
REF_EXE = os.path.join(REF_DIR, "MC_ref")
REF_EXE_PATCHED = os.path.join(REF_DIR, "MC_ref_patched")
XS_DIR = os.path.join(REF_DIR, "xs_library")

RNG_GLOBAL, RNG_HISTORY = 0, 1
PICK_CDF, PICK_FLOOR = 0, 1


class OracleCycle(C.Structure):
    _fields_ = [(n, C.c_double) for n in ("k_cycle", "k_avg", "k_uncer", "H", "k_sum_C", "k_sum_TL", "k_sq_C",
                                          "k_sq_TL", "H_sum")] + \
               [(n, C.c_uint64) for n in ("n_sites", "n_tracks", "n_collisions", "n_histories", "n_draws")]

    def as_dict(self):
        return {f: getattr(self, f) for f, _ in self._fields_}


_oracle = None


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def oracle():
    global _oracle
    if _oracle is None:
        L = C.CDLL(ORACLE_LIB)
        vp, i32, i64, u64, d = C.c_void_p, C.c_int32, C.c_int64, C.c_uint64, C.c_double
        L.mco_create.restype = vp
        L.mco_create.argtypes = [vp, C.c_int, C.c_int]
        L.mco_destroy.argtypes = [vp]
        L.mco_set_shard.argtypes = [vp, u64, u64]
        L.mco_run_cycle.argtypes = [vp, C.POINTER(OracleCycle)]
        L.mco_transport_cycle.argtypes = [vp]
        L.mco_get_partials.argtypes = [vp, vp, vp]
        L.mco_get_history_k.restype = i64
        L.mco_get_history_k.argtypes = [vp, vp, vp]
        L.mco_bank_size.restype = i64
        L.mco_bank_size.argtypes = [vp]
        L.mco_get_bank.argtypes = [vp, vp, vp]
        L.mco_set_source_bank.argtypes = [vp, vp, vp, i64]
        L.mco_get_tally_partials.argtypes = [vp, vp, vp]
        L.mco_close_cycle.argtypes = [vp, vp, vp, vp, vp, C.POINTER(OracleCycle)]
        L.mco_end_simulation.argtypes = [vp]
        L.mco_get_tallies.argtypes = [vp, vp, vp]
        L.mco_get_k.restype = d
        L.mco_get_k.argtypes = [vp]
        L.mco_get_seed.restype = u64
        L.mco_get_seed.argtypes = [vp]
        L.mco_xs_lookup.argtypes = [vp, C.c_int, vp, i64, vp]
        L.mco_select_channel.argtypes = [vp, C.c_int, C.c_int, vp, vp, i64, vp]
        L.mco_beta.argtypes = [vp, C.c_int, vp, i64, vp]
        L.mco_binary_search.argtypes = [d, vp, C.c_int]
        L.mco_interpolate.restype = d
        L.mco_interpolate.argtypes = [d] * 5
        L.mco_geometry_quad.restype = d
        L.mco_geometry_quad.argtypes = [d] * 3
        L.mco_scatter_direction.argtypes = [vp, d, d, vp]
        L.mco_lcg_next.restype = u64
        L.mco_lcg_next.argtypes = [u64]
        L.mco_lcg_skip.restype = u64
        L.mco_lcg_skip.argtypes = [u64, u64]
        L.mco_surface_eval.restype = d
        L.mco_surface_eval.argtypes = [vp, C.c_int, vp]
        L.mco_surface_distance.restype = d
        L.mco_surface_distance.argtypes = [vp, C.c_int, vp, vp]
        L.mco_surface_reflect.argtypes = [vp, C.c_int, vp]
        L.mco_search_cell.argtypes = [vp, vp]
        L.mco_surface_intersect.argtypes = [vp, C.c_int, vp, vp, vp]
        L.mco_scatter_sample.argtypes = [vp, C.c_int, vp, vp]
        L.mco_watt_sample.restype = d
        L.mco_watt_sample.argtypes = [vp, C.c_int, vp, d]
        L.mco_speed_of_energy.restype = d
        L.mco_speed_of_energy.argtypes = [d]
        L.mco_energy_of_speed.restype = d
        L.mco_energy_of_speed.argtypes = [d]
        _oracle = L
    return _oracle


class Oracle:
    """One oracle run over a mc_old_b200.Deck."""

    def __init__(self, deck, rng_mode=RNG_GLOBAL, pick_mode=PICK_CDF):
        self.deck = deck
        self.L = oracle()
        self.h = self.L.mco_create(deck.problem, rng_mode, pick_mode)
        self.nt = deck.info["n_tallies"]

    def close(self):
        if self.h:
            self.L.mco_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_shard(self, begin, count):
        self.L.mco_set_shard(self.h, begin, count)

    def run_cycle(self):
        r = OracleCycle()
        if self.L.mco_run_cycle(self.h, C.byref(r)) != 0:
            raise RuntimeError("oracle: particle lost / empty source bank")
        return r

    def run_cycle_keep_bank(self):
        """run_cycle through the split phases, returning (result, fission sites, cells) of the generation"""
        self.transport_cycle()
        sums, counts = self.partials()
        sites, cells = self.bank()
        ts, tq = self.tally_partials()
        r = self.close_cycle(sums, counts, ts, tq)
        if self.deck.info["ksearch"]:
            self.set_source_bank(sites, cells)
        return r, sites, cells

    def run(self):
        info = self.deck.info
        res = [self.run_cycle() for _ in range(info["n_cycle"])]
        self.L.mco_end_simulation(self.h)
        return res

    # split phases (sharded runs)
    def transport_cycle(self):
        if self.L.mco_transport_cycle(self.h) != 0:
            raise RuntimeError("oracle: particle lost / empty source bank")

    def partials(self):
        s = np.zeros(5); n = np.zeros(4, dtype=np.uint64)
        self.L.mco_get_partials(self.h, _p(s), _p(n))
        return s, n

    def history_k(self, n):
        kC = np.zeros(max(n, 1)); kTL = np.zeros(max(n, 1))
        got = self.L.mco_get_history_k(self.h, _p(kC), _p(kTL))
        return kC[:got], kTL[:got]

    def bank(self):
        n = self.L.mco_bank_size(self.h)
        sites = np.zeros((max(n, 1), 8)); cells = np.zeros(max(n, 1), dtype=np.int32)
        self.L.mco_get_bank(self.h, _p(sites), _p(cells))
        return sites[:n], cells[:n]

    def set_source_bank(self, sites, cells):
        sites = np.ascontiguousarray(sites, dtype=np.float64); cells = np.ascontiguousarray(cells, dtype=np.int32)
        self.L.mco_set_source_bank(self.h, _p(sites), _p(cells), sites.shape[0])

    def tally_partials(self):
        s = np.zeros(max(self.nt, 1)); q = np.zeros(max(self.nt, 1))
        self.L.mco_get_tally_partials(self.h, _p(s), _p(q))
        return s, q

    def close_cycle(self, sums, counts, tsum, tsq):
        r = OracleCycle()
        sums = np.ascontiguousarray(sums, dtype=np.float64); counts = np.ascontiguousarray(counts, dtype=np.uint64)
        tsum = np.ascontiguousarray(tsum, dtype=np.float64); tsq = np.ascontiguousarray(tsq, dtype=np.float64)
        self.L.mco_close_cycle(self.h, _p(sums), _p(counts), _p(tsum), _p(tsq), C.byref(r))
        return r

    def end_simulation(self):
        self.L.mco_end_simulation(self.h)

    def tallies(self):
        m = np.zeros(max(self.nt, 1)); u = np.zeros(max(self.nt, 1))
        self.L.mco_get_tallies(self.h, _p(m), _p(u))
        return m[:self.nt], u[:self.nt]

    @property
    def k(self):
        return self.L.mco_get_k(self.h)


def xs_lookup(deck, material, E):
    E = np.ascontiguousarray(E, dtype=np.float64)
    out = np.empty((E.size, 5))
    oracle().mco_xs_lookup(deck.problem, material, _p(E), E.size, _p(out))
    return out


def select_channel(deck, material, kind, E, xi):
    E = np.ascontiguousarray(E, dtype=np.float64); xi = np.ascontiguousarray(xi, dtype=np.float64)
    out = np.empty(E.size, dtype=np.int32)
    oracle().mco_select_channel(deck.problem, material, kind, _p(E), _p(xi), E.size, _p(out))
    return out


def beta(deck, nuclide, E):
    E = np.ascontiguousarray(E, dtype=np.float64)
    out = np.empty(E.size)
    oracle().mco_beta(deck.problem, nuclide, _p(E), E.size, _p(out))
    return out


# ---------------------------------------------------------------------------------------------
# the compiled reference
# ---------------------------------------------------------------------------------------------
def have_ref():
    return os.path.exists(REF_EXE) and os.path.exists(REF_HARNESS) and os.path.isdir(XS_DIR)


def parse_ref_output(path):
    """Parse the text `output.h5` written by oracle/h5stub/H5Cpp.h."""
    out = {}
    with open(path) as f:
        for line in f:
            if not line.startswith("D "):
                continue
            head, _, vals = line[2:].partition(" :")
            # dataset paths may contain spaces ("Estimator #1"): type token is the first of f64|u64|str after the path
            m = re.match(r"(.*) (f64|u64|str) (\d+)((?: \d+)*)$", head)
            name, typ = m.group(1), m.group(2)
            if typ == "f64":
                out[name] = np.array([float(v) for v in vals.split()])
            elif typ == "u64":
                out[name] = np.array([int(v) for v in vals.split()], dtype=np.uint64)
            else:
                out[name] = vals.strip()
    return out


def run_ref(deck_dir, patched=False, timeout=600):
    """Run the compiled reference on <deck_dir>/input.xml (CWD = oracle/_ref so ./xs_library resolves).
    Returns (stdout, parsed output)."""
    exe = REF_EXE_PATCHED if patched else REF_EXE
    r = subprocess.run([exe, deck_dir], cwd=REF_DIR, capture_output=True, text=True, timeout=timeout)
    if r.returncode != 0:
        raise RuntimeError("reference failed (%d): %s" % (r.returncode, r.stdout[-500:] + r.stderr[-500:]))
    return r.stdout, parse_ref_output(os.path.join(deck_dir, "output.h5"))


_harness = None


def harness():
    global _harness
    if _harness is None:
        L = C.CDLL(REF_HARNESS)
        vp, d, u64 = C.c_void_p, C.c_double, C.c_uint64
        L.refh_set_seed.argtypes = [u64]
        L.refh_get_seed.restype = u64
        L.refh_draws.restype = u64
        L.refh_urand.restype = d
        L.refh_inject.argtypes = [vp, C.c_int]
        L.refh_create.restype = vp
        L.refh_create.argtypes = [C.c_char_p]
        L.refh_destroy.argtypes = [vp]
        L.refh_counts.argtypes = [vp, vp]
        L.refh_sigma.argtypes = [vp, C.c_int, vp, C.c_int, vp]
        L.refh_micro.argtypes = [vp, C.c_int, vp, C.c_int, vp]
        L.refh_select.argtypes = [vp, C.c_int, C.c_int, vp, vp, C.c_int, vp]
        L.refh_binary_search.argtypes = [d, vp, C.c_int]
        L.refh_interpolate.restype = d
        L.refh_interpolate.argtypes = [d] * 5
        L.refh_geometry_quad.restype = d
        L.refh_geometry_quad.argtypes = [d] * 3
        L.refh_scatter_direction.argtypes = [vp, d, vp]
        L.refh_surface.argtypes = [vp, C.c_int, vp, vp, vp]
        L.refh_scatter_sample.argtypes = [vp, C.c_int, vp]
        L.refh_watt.restype = d
        L.refh_watt.argtypes = [vp, C.c_int, d]
        L.refh_isotropic.argtypes = [vp]
        L.refh_particle_speed.argtypes = [vp, d, vp]
        _harness = L
    return _harness


class RefSim:
    """The reference's Simulator object, constructed in-process from a deck directory."""

    def __init__(self, deck_dir):
        self.L = harness()
        cwd = os.getcwd()
        os.chdir(REF_DIR)  # ./xs_library is resolved relative to the CWD (setup.cpp:326)
        try:
            d = deck_dir if deck_dir.endswith("/") else deck_dir + "/"
            self.h = self.L.refh_create(d.encode())
        finally:
            os.chdir(cwd)

    def sigma(self, mat, E):
        E = np.ascontiguousarray(E, dtype=np.float64)
        out = np.empty((E.size, 6))
        self.L.refh_sigma(self.h, mat, _p(E), E.size, _p(out))
        return out

    def micro(self, nuc, E):
        E = np.ascontiguousarray(E, dtype=np.float64)
        out = np.empty((E.size, 6))
        self.L.refh_micro(self.h, nuc, _p(E), E.size, _p(out))
        return out

    def select(self, mat, kind, E, xi):
        E = np.ascontiguousarray(E, dtype=np.float64); xi = np.ascontiguousarray(xi, dtype=np.float64)
        out = np.empty(E.size, dtype=np.int32)
        self.L.refh_select(self.h, mat, kind, _p(E), _p(xi), E.size, _p(out))
        return out

    def surface(self, s, pos, dir):
        pos = np.ascontiguousarray(pos, dtype=np.float64); dir = np.ascontiguousarray(dir, dtype=np.float64)
        out = np.empty(6)
        self.L.refh_surface(self.h, s, _p(pos), _p(dir), _p(out))
        return out

    def scatter_sample(self, nuc, dir, E, seed):
        io = np.array([dir[0], dir[1], dir[2], E, 0.0])
        self.L.refh_set_seed(seed)
        self.L.refh_scatter_sample(self.h, nuc, _p(io))
        return io, self.L.refh_get_seed()

    def watt(self, nuc, E, seed):
        self.L.refh_set_seed(seed)
        v = self.L.refh_watt(self.h, nuc, E)
        return v, self.L.refh_get_seed()
