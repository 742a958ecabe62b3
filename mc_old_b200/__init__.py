"""mc_old_b200 — Python face of the B200-native transport loop.

Thin ctypes bindings over the two shared libraries of the product:

* ``libmcbhost.so`` (include/mcb200_host.h): input.xml + xs_library -> flattened
  ``mcb_problem`` (replaces the reference's ``Simulator`` constructor,
  src/simulator/setup.cpp:30-1069).  Pure host code, no GPU needed.
* ``libmcb200.so`` (include/mcb200.h): the CUDA transport loop behind the C-ABI
  (replaces ``Simulator::start()``, src/simulator/handler.cpp:11-48).

There is no CPU fallback: anything that computes raises ``RuntimeError`` when the
CUDA library is missing or no device is present.  PyTorch is not imported here.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
REPO_ROOT = os.path.dirname(_HERE)
HOST_LIB = os.path.join(_HERE, "libmcbhost.so")
CUDA_LIB = os.environ.get("MCB200_LIB") or os.path.join(_HERE, "libmcb200.so")

IGNORE_TRMM = 1


def default_xs_dir() -> str:
    """Directory holding <ZAID>.txt.  The reference reads ./xs_library relative to the CWD
    (setup.cpp:326); MCB_XS_LIBRARY overrides; data/xs_library is where build() puts a copy of the reference's
    cross-section text files (data, git-ignored) when the reference tree is present."""
    for cand in (os.environ.get("MCB_XS_LIBRARY"), os.path.join(os.getcwd(), "xs_library"),
                 os.path.join(REPO_ROOT, "data", "xs_library")):
        if cand and os.path.isdir(cand):
            return cand
    raise FileNotFoundError("xs_library not found (set MCB_XS_LIBRARY)")


class CycleResult(C.Structure):
    _fields_ = [("k_cycle", C.c_double), ("k_avg", C.c_double), ("k_uncer", C.c_double), ("H", C.c_double),
                ("H_cycle_conventional", C.c_double),
                ("k_sum_C", C.c_double), ("k_sum_TL", C.c_double), ("k_sq_C", C.c_double), ("k_sq_TL", C.c_double),
                ("n_sites", C.c_uint64), ("n_histories", C.c_uint64), ("n_tracks", C.c_uint64),
                ("n_collisions", C.c_uint64), ("n_lookups", C.c_uint64), ("n_crossings", C.c_uint64),
                ("ms_transport", C.c_double), ("ms_exchange", C.c_double),
                ("n_iterations", C.c_int32), ("lost", C.c_int32), ("n_kernel_launches", C.c_uint64)]

    def as_dict(self):
        return {f: getattr(self, f) for f, _ in self._fields_}


class StageTimes(C.Structure):
    _fields_ = [(n, C.c_double) for n in ("ms_source", "ms_lookup", "ms_flight", "ms_cross", "ms_collide",
                                          "ms_closeout", "ms_bank")] + \
               [(n, C.c_uint64) for n in ("n_source", "n_lookup", "n_flight", "n_cross", "n_collide", "n_closeout",
                                          "n_bank", "units_lookup")] + \
               [("ms_finish", C.c_double), ("n_finish", C.c_uint64), ("ms_step", C.c_double), ("n_step", C.c_uint64)]

    def as_dict(self):
        return {f: getattr(self, f) for f, _ in self._fields_}


class Config(C.Structure):
    _fields_ = [("device", C.c_int32), ("rank", C.c_int32), ("world", C.c_int32), ("reserved", C.c_int32),
                ("bank_capacity", C.c_int64), ("site_capacity", C.c_int64), ("stream", C.c_void_p)]


_host = None
_cuda = None


def host_lib():
    global _host
    if _host is None:
        if not os.path.exists(HOST_LIB):
            raise RuntimeError("libmcbhost.so is not built: run `python -c 'import __graft_entry__ as g; g.build()'`")
        L = C.CDLL(HOST_LIB)
        L.mcbh_load_deck.restype = C.c_void_p
        L.mcbh_load_deck.argtypes = [C.c_char_p, C.c_char_p, C.c_int]
        L.mcbh_load_deck_string.restype = C.c_void_p
        L.mcbh_load_deck_string.argtypes = [C.c_char_p, C.c_char_p, C.c_int]
        L.mcbh_free_deck.argtypes = [C.c_void_p]
        L.mcbh_last_error.restype = C.c_char_p
        L.mcbh_problem.restype = C.c_void_p
        L.mcbh_problem.argtypes = [C.c_void_p]
        L.mcbh_set_run.argtypes = [C.c_void_p, C.c_uint64, C.c_uint64, C.c_uint64, C.c_uint64]
        L.mcbh_info.argtypes = [C.c_void_p, C.POINTER(C.c_int64)]
        L.mcbh_name.restype = C.c_char_p
        L.mcbh_name.argtypes = [C.c_void_p, C.c_int, C.c_int]
        L.mcbh_mode.restype = C.c_char_p
        L.mcbh_mode.argtypes = [C.c_void_p]
        L.mcbh_search_cell.argtypes = [C.c_void_p, C.c_double, C.c_double, C.c_double]
        L.mcbh_estimator_info.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_int64)]
        L.mcbh_filter_info.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_int64)]
        L.mcbh_filter_grid.restype = C.POINTER(C.c_double)
        L.mcbh_filter_grid.argtypes = [C.c_void_p]
        L.mcbh_simulation_name.restype = C.c_char_p
        L.mcbh_simulation_name.argtypes = [C.c_void_p]
        _host = L
    return _host


def cuda_lib():
    """The CUDA library; raises when it is missing — there is no fallback path."""
    global _cuda
    if _cuda is None:
        if not os.path.exists(CUDA_LIB):
            raise RuntimeError("libmcb200.so (CUDA kernels) is not built; there is no CPU fallback")
        L = C.CDLL(CUDA_LIB)
        vp, i32, i64, u64, dp = C.c_void_p, C.c_int32, C.c_int64, C.c_uint64, C.POINTER(C.c_double)
        L.mcb_create.argtypes = [vp, C.POINTER(Config), C.POINTER(vp)]
        L.mcb_destroy.argtypes = [vp]
        L.mcb_last_error.restype = C.c_char_p
        L.mcb_last_error.argtypes = [vp]
        L.mcb_comm_unique_id.argtypes = [C.c_char_p]
        L.mcb_comm_init.argtypes = [vp, C.c_char_p]
        L.mcb_run_cycle.argtypes = [vp, C.POINTER(CycleResult)]
        L.mcb_get_tallies.argtypes = [vp, vp, vp, i64]
        L.mcb_get_k.restype = C.c_double
        L.mcb_get_k.argtypes = [vp]
        L.mcb_set_k.argtypes = [vp, C.c_double]
        L.mcb_get_stage_times.argtypes = [vp, C.POINTER(StageTimes)]
        L.mcb_reset_stage_times.argtypes = [vp]
        L.mcb_set_stage_timing.argtypes = [vp, C.c_int]
        L.mcb_get_fission_bank.restype = i64
        L.mcb_get_fission_bank.argtypes = [vp, vp, vp, i64]
        L.mcb_set_source_bank.argtypes = [vp, vp, vp, i64]
        L.mcb_get_source_bank.restype = i64
        L.mcb_get_source_bank.argtypes = [vp, vp, vp, i64]
        L.mcb_get_history_k.restype = i64
        L.mcb_get_history_k.argtypes = [vp, vp, vp, i64]
        L.mcb_beta_batch.argtypes = [vp, i32, i32, vp, i64, vp]
        L.mcb_device_count.restype = C.c_int
        L.mcb_xs_lookup_batch.argtypes = [vp, i32, vp, i64, vp]
        L.mcb_xs_lookup_device.argtypes = [vp, i32, vp, i64, vp, C.POINTER(C.c_float)]
        L.mcb_run_cycle_host.argtypes = [vp, vp, vp, i64, vp, vp, i64, C.POINTER(i64), vp]
        L.mcb_select_channel_batch.argtypes = [vp, i32, i32, vp, vp, i64, vp]
        L.mcb_rng_batch.argtypes = [vp, vp, i64, i32, vp]
        L.mcb_geometry_batch.argtypes = [vp, vp, vp, vp, i64, vp]
        L.mcb_search_cell_batch.argtypes = [vp, vp, i64, vp]
        L.mcb_scatter_batch.argtypes = [vp, i32, vp, i64, vp]
        L.mcb_watt_batch.argtypes = [vp, i32, vp, vp, i64, vp]
        L.mcb_division_batch.argtypes = [vp, vp, vp, i64, vp, vp]
        L.mcb_walk_launch_info.argtypes = [vp, i32, vp]
        L.mcb_shard_range.argtypes = [u64, i32, i32, C.POINTER(u64), C.POINTER(u64)]
        L.mcb_shard_range.restype = None
        _cuda = L
    return _cuda


def _ptr(a: np.ndarray):
    return a.ctypes.data_as(C.c_void_p)


class Deck:
    """Problem setup: mirrors ``Simulator(io_dir)`` (setup.cpp:30).  ``Deck(path)`` loads ``path/input.xml``;
    ``Deck(xml=text)`` parses a string."""

    INFO = ("n_sample", "n_cycle", "n_passive", "ksearch", "n_nuclides", "n_materials", "n_surfaces", "n_cells",
            "n_estimators", "n_tallies", "n_sources", "entropy_on", "n_xs_rows", "n_scores", "n_filters",
            "trmm_present")

    def __init__(self, io_dir: Optional[str] = None, xml: Optional[str] = None, xs_dir: Optional[str] = None,
                 flags: int = 0):
        L = host_lib()
        xs = (xs_dir or default_xs_dir()).encode()
        if xml is not None:
            h = L.mcbh_load_deck_string(xml.encode(), xs, flags)
        else:
            h = L.mcbh_load_deck(str(io_dir).encode(), xs, flags)
        if not h:
            raise ValueError(L.mcbh_last_error().decode())
        self._h = h

    def close(self):
        if getattr(self, "_h", None):
            host_lib().mcbh_free_deck(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def problem(self) -> int:
        """const mcb_problem* as an integer address."""
        return host_lib().mcbh_problem(self._h)

    def set_run(self, n_sample=0, n_cycle=0, n_passive=0, seed=0):
        host_lib().mcbh_set_run(self._h, int(n_sample), int(n_cycle), int(n_passive), int(seed))

    @property
    def info(self) -> dict:
        out = (C.c_int64 * 16)()
        host_lib().mcbh_info(self._h, out)
        return dict(zip(self.INFO, list(out)))

    def name(self, kind: int, index: int) -> Optional[str]:
        s = host_lib().mcbh_name(self._h, kind, index)
        return s.decode() if s is not None else None

    def estimators(self) -> list:
        """Tally layout for reporting (Estimator.cpp:280-295,368-422): one dict per estimator with its scores and
        filters; tally t of score k lives at tally_begin + k*prod(filter sizes) + row-major filter index."""
        L = host_lib()
        grid = L.mcbh_filter_grid(self._h)
        out = []
        for e in range(self.info["n_estimators"]):
            v = (C.c_int64 * 8)()
            L.mcbh_estimator_info(self._h, e, v)
            filters = []
            for f in range(v[3], v[3] + v[4]):
                w = (C.c_int64 * 4)()
                L.mcbh_filter_info(self._h, f, w)
                filters.append({"type": int(w[0]), "grid": [grid[i] for i in range(w[1], w[1] + w[2])], "size": int(w[3])})
            out.append({"name": self.name(4, e), "attach": int(v[0]),
                        "scores": [self.name(5, k) for k in range(v[1], v[1] + v[2])], "filters": filters,
                        "tally_begin": int(v[5]), "n_tallies": int(v[6]), "simulate": int(v[7])})
        return out

    @property
    def simulation_name(self) -> str:
        return host_lib().mcbh_simulation_name(self._h).decode()

    @property
    def mode(self) -> str:
        return host_lib().mcbh_mode(self._h).decode()

    def tdmc(self):
        """(census times, interval lengths as the reference computes them) of a time-dependent deck, else two empty arrays"""
        L = host_lib()
        L.mcbh_tdmc.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        n = L.mcbh_tdmc(self._h, None, None)
        t = np.zeros(n); dt = np.zeros(n)
        if n:
            L.mcbh_tdmc(self._h, t.ctypes.data, dt.ctypes.data)
        return t, dt

    def search_cell(self, x, y, z) -> int:
        return host_lib().mcbh_search_cell(self._h, x, y, z)


class Context:
    """Device context of one rank (mcb_ctx): owns tables, banks and tallies on one GPU."""

    def __init__(self, deck: Deck, device: int = 0, rank: int = 0, world: int = 1, bank_capacity: int = 0,
                 site_capacity: int = 0, stream: int = 0, stage_times: bool = False, split_stages: bool = False):
        L = cuda_lib()
        self.deck = deck
        cfg = Config(device, rank, world, (1 if stage_times else 0) | (2 if split_stages else 0), bank_capacity, site_capacity, stream or None)
        h = C.c_void_p()
        rc = L.mcb_create(deck.problem, C.byref(cfg), C.byref(h))
        if rc != 0:
            raise RuntimeError("mcb_create failed (%d): %s" % (rc, L.mcb_last_error(None).decode()))
        self._h = h

    def close(self):
        if getattr(self, "_h", None):
            cuda_lib().mcb_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc):
        if rc != 0:
            raise RuntimeError("mcb200 error %d: %s" % (rc, cuda_lib().mcb_last_error(self._h).decode()))

    # -- multi-GPU plumbing --
    @staticmethod
    def unique_id() -> bytes:
        buf = C.create_string_buffer(128)
        rc = cuda_lib().mcb_comm_unique_id(buf)
        if rc != 0:
            raise RuntimeError("mcb_comm_unique_id failed")
        return buf.raw

    def comm_init(self, uid: bytes):
        self._check(cuda_lib().mcb_comm_init(self._h, uid))

    # -- transport --
    def run_cycle(self) -> CycleResult:
        r = CycleResult()
        self._check(cuda_lib().mcb_run_cycle(self._h, C.byref(r)))
        return r

    def tallies(self):
        n = self.deck.info["n_tallies"]
        mean = np.zeros(max(n, 1)); uncer = np.zeros(max(n, 1))
        self._check(cuda_lib().mcb_get_tallies(self._h, _ptr(mean), _ptr(uncer), n))
        return mean[:n], uncer[:n]

    @property
    def k(self) -> float:
        return cuda_lib().mcb_get_k(self._h)

    @k.setter
    def k(self, v: float):
        cuda_lib().mcb_set_k(self._h, float(v))

    def stage_times(self) -> dict:
        s = StageTimes()
        self._check(cuda_lib().mcb_get_stage_times(self._h, C.byref(s)))
        return s.as_dict()

    def set_stage_timing(self, on: bool):
        cuda_lib().mcb_set_stage_timing(self._h, 1 if on else 0)

    def reset_stage_times(self):
        cuda_lib().mcb_reset_stage_times(self._h)

    def fission_bank(self, max_n: int):
        sites = np.zeros((max(max_n, 1), 8)); cells = np.zeros(max(max_n, 1), dtype=np.int32)
        n = cuda_lib().mcb_get_fission_bank(self._h, _ptr(sites), _ptr(cells), max_n)
        if n < 0:
            self._check(int(n))
        return sites[:n], cells[:n]

    def source_bank(self, max_n: int, sites: Optional[np.ndarray] = None, cells: Optional[np.ndarray] = None):
        """Global source bank of the next cycle; pass preallocated (pinned) arrays to avoid allocations."""
        if sites is None:
            sites = np.zeros((max(max_n, 1), 8)); cells = np.zeros(max(max_n, 1), dtype=np.int32)
        n = cuda_lib().mcb_get_source_bank(self._h, _ptr(sites), _ptr(cells), max_n)
        if n < 0:
            self._check(int(n))
        return sites[:n], cells[:n]

    def set_source_bank(self, sites: np.ndarray, cells: np.ndarray):
        sites = np.ascontiguousarray(sites, dtype=np.float64).reshape(-1, 8)
        cells = np.ascontiguousarray(cells, dtype=np.int32)
        self._check(cuda_lib().mcb_set_source_bank(self._h, _ptr(sites), _ptr(cells), sites.shape[0]))

    def run_cycle_host(self, sites_in: np.ndarray, cells_in: np.ndarray, sites_out: np.ndarray, cells_out: np.ndarray):
        """One generation with the source bank in host arrays on both sides (pinned arrays make the copies
        asynchronous); returns (CycleResult, new bank sites view, cells view)."""
        r = CycleResult()
        n_out = C.c_int64()
        sites_in = np.ascontiguousarray(sites_in, dtype=np.float64).reshape(-1, 8)
        cells_in = np.ascontiguousarray(cells_in, dtype=np.int32)
        self._check(cuda_lib().mcb_run_cycle_host(self._h, _ptr(sites_in), _ptr(cells_in), sites_in.shape[0], _ptr(sites_out),
                                                  _ptr(cells_out), sites_out.shape[0], C.byref(n_out), C.byref(r)))
        return r, sites_out[:n_out.value], cells_out[:n_out.value]

    def history_k(self, n: int):
        kC = np.zeros(max(n, 1)); kTL = np.zeros(max(n, 1))
        got = cuda_lib().mcb_get_history_k(self._h, _ptr(kC), _ptr(kTL), n)
        if got < 0:
            self._check(int(got))
        return kC[:got], kTL[:got]

    def beta(self, material: int, local_nuclide: int, E: np.ndarray) -> np.ndarray:
        E = np.ascontiguousarray(E, dtype=np.float64)
        out = np.empty(E.size)
        self._check(cuda_lib().mcb_beta_batch(self._h, material, local_nuclide, _ptr(E), E.size, _ptr(out)))
        return out

    # -- parity / bench entry points (host buffers) --
    def xs_lookup(self, material: int, E: np.ndarray) -> np.ndarray:
        E = np.ascontiguousarray(E, dtype=np.float64)
        out = np.empty((E.size, 5))
        self._check(cuda_lib().mcb_xs_lookup_batch(self._h, material, _ptr(E), E.size, _ptr(out)))
        return out

    def xs_lookup_device(self, material: int, dE_ptr: int, n: int, dout_ptr: int) -> float:
        ms = C.c_float()
        self._check(cuda_lib().mcb_xs_lookup_device(self._h, material, dE_ptr, n, dout_ptr, C.byref(ms)))
        return ms.value

    def select_channel(self, material: int, kind: int, E: np.ndarray, xi: np.ndarray) -> np.ndarray:
        E = np.ascontiguousarray(E, dtype=np.float64); xi = np.ascontiguousarray(xi, dtype=np.float64)
        out = np.empty(E.size, dtype=np.int32)
        self._check(cuda_lib().mcb_select_channel_batch(self._h, material, kind, _ptr(E), _ptr(xi), E.size, _ptr(out)))
        return out

    def rng(self, nps: np.ndarray, ndraw: int) -> np.ndarray:
        nps = np.ascontiguousarray(nps, dtype=np.uint64)
        out = np.empty((nps.size, ndraw), dtype=np.uint64)
        self._check(cuda_lib().mcb_rng_batch(self._h, _ptr(nps), nps.size, ndraw, _ptr(out)))
        return out

    def geometry(self, cell: np.ndarray, pos: np.ndarray, dir: np.ndarray) -> np.ndarray:
        cell = np.ascontiguousarray(cell, dtype=np.int32)
        pos = np.ascontiguousarray(pos, dtype=np.float64); dir = np.ascontiguousarray(dir, dtype=np.float64)
        out = np.empty((cell.size, 3))
        self._check(cuda_lib().mcb_geometry_batch(self._h, _ptr(cell), _ptr(pos), _ptr(dir), cell.size, _ptr(out)))
        return out

    def search_cell(self, pos: np.ndarray) -> np.ndarray:
        pos = np.ascontiguousarray(pos, dtype=np.float64)
        out = np.empty(pos.shape[0], dtype=np.int32)
        self._check(cuda_lib().mcb_search_cell_batch(self._h, _ptr(pos), pos.shape[0], _ptr(out)))
        return out

    def scatter(self, nuclide: int, nps: np.ndarray, io5: np.ndarray) -> np.ndarray:
        nps = np.ascontiguousarray(nps, dtype=np.uint64)
        io = np.array(io5, dtype=np.float64, order="C", copy=True)
        self._check(cuda_lib().mcb_scatter_batch(self._h, nuclide, _ptr(nps), nps.size, _ptr(io)))
        return io

    def watt(self, nuclide: int, nps: np.ndarray, E: np.ndarray) -> np.ndarray:
        nps = np.ascontiguousarray(nps, dtype=np.uint64); E = np.ascontiguousarray(E, dtype=np.float64)
        out = np.empty(E.size)
        self._check(cuda_lib().mcb_watt_batch(self._h, nuclide, _ptr(nps), _ptr(E), E.size, _ptr(out)))
        return out


    def walk_launch_info(self, scoring=False):
        """{registers, grid, block, smem} of the walk kernel as launched on this device"""
        out = np.zeros(4, dtype=np.int32)
        self._check(cuda_lib().mcb_walk_launch_info(self._h, 1 if scoring else 0, _ptr(out)))
        return dict(zip(("registers", "grid", "block", "smem"), (int(x) for x in out)))

    def division(self, a: np.ndarray, b: np.ndarray):
        """a / b through the kernels' shared-reciprocal form and as the plain IEEE division (mcb_division_batch)"""
        a = np.ascontiguousarray(a, dtype=np.float64); b = np.ascontiguousarray(b, dtype=np.float64)
        o1, o2 = np.empty(a.size), np.empty(a.size)
        self._check(cuda_lib().mcb_division_batch(self._h, _ptr(a), _ptr(b), a.size, _ptr(o1), _ptr(o2)))
        return o1, o2


def shard_range(n: int, rank: int, world: int):
    """History slice owned by `rank`; pure host arithmetic, identical on every rank."""
    b, c = C.c_uint64(), C.c_uint64()
    cuda_lib().mcb_shard_range(n, rank, world, C.byref(b), C.byref(c))
    return b.value, c.value
