#!/usr/bin/env python3
"""bench.py — histories/s of the B200 transport loop on the HEU-sphere k-eigenvalue problem.

    python bench.py --gpus N --steps K --warmup W            # our arm (CUDA, one rank per GPU)
    python bench.py --impl reference --gpus N --steps K --warmup W   # the reference's CPU implementation, rank 0 only

Workload (BASELINE.json north_star target): examples/HEU_sphere_criticality physics (bare HEU sphere r = 7.68 cm,
U-235 + U-238, k-eigenvalue power iteration), scaled to --samples histories per generation PER GPU (weak scaling;
default 1e7, i.e. 8e7 ~ the 1e8 target at 8 GPUs).  A "step" is one generation: source resampling from the fission
bank, the event loop until every history is finished, fission-bank ordering, k close-out and, for N > 1, the NCCL
all-gather of the ranks' close-out sums; the fission bank stays where it was banked and the next generation
reads the sites it draws in place over NVLink (CUDA IPC peer mappings, draws sorted by site index).

`value`  histories/s with the source bank resident in HBM (device time, CUDA events on the launch stream, max over
         ranks).
`e2e`    the same step driven through the C-ABI with HOST buffers (mcb_run_cycle_host): every step takes the global
         source bank from pinned host memory, runs the generation, and puts the new global bank back into pinned host
         memory.
`roofline`  the dominant stage kernel of the timed loop: algorithmic bytes per launch / mean launch time (CUDA events
         per launch in a separate pass, stage timing costs a few percent so it is off in the timed region).
`cpu_baseline`  the compiled reference (oracle/_ref/MC_ref, single-threaded by construction) on the same deck at a
         bounded history count, timed per generation from its own stdout on this box's host cores.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOAD = "HEU_sphere_criticality"
# Results must not depend on the number of GPUs (per-history streams, canonical bank order, exact integer sums): two
# generations of the HEU sphere at 4e4 histories, k per generation as hex and the fission-bank size, recorded from a
# 1-GPU run (tests/golden/g_independence.json, written by `bench.py --record-g-independence`); every run at any GPU
# count repeats them and must reproduce them bit for bit.
G_IND_FILE = os.path.join(ROOT, "tests", "golden", "g_independence.json")
G_IND_SAMPLES, G_IND_CYCLES = 40000, 2
# algorithmic bytes per unit of each stage kernel; Nn = nuclides of the material (HEU: 2)
BYTES_LOOKUP = lambda nn: 72 + 100 * nn   # E + mat 12, hash 4, bracket 16, Nn x (idx 4 + 2 rows x 48), 5 Sigma out 40
BYTES_FLIGHT = 180                        # queue 4 + state in 112 + state out 44 + k_TL rmw 16 + event queue 4
BYTES_COLLIDE = lambda nn: 220 + 2 * 96 * nn + 27  # state in 144 + out 76 + 2 selections x Nn x 2 rows + 0.35 sites x 76
BYTES_CROSS = 140
# walk kernel: the whole-loop figure of SURVEY §8d per track = particle record in + out (224) + one xs lookup
# (72 + 100 Nn) + 76-byte fission sites (one per ~3.8 tracks in HEU)
BYTES_STEP = lambda nn: 224 + 72 + 100 * nn + 20


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured"
    except Exception:
        return 6650.0, "fallback"


def ncu_traffic(kernel, launch=None):
    """dram__bytes_read.sum + dram__bytes_write.sum of one launch of `kernel` from the newest committed
    `ncu --set full` summary under profiles/ (same workload: 1e7 histories per generation).  A capture whose launch
    shape (registers per thread, grid) differs from the kernel that is running is stale and is refused: returns
    (None, path, why).  None when there is no capture."""
    import glob
    import re
    best = None
    for path in sorted(glob.glob(os.path.join(ROOT, "profiles", "r*_%s_ncu_full_summary.txt" % kernel))):
        tot, seen, regs, grid = 0.0, 0, None, None
        for line in open(path):
            m = re.match(r"\s*dram__bytes_(read|write)\.sum\s+([0-9.]+)\s+(\w+)", line)
            if m:
                tot += float(m.group(2)) * {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(m.group(3), 0.0)
                seen += 1
            m = re.match(r"\s*launch__registers_per_thread\s+([0-9]+)", line)
            if m:
                regs = int(m.group(1))
            m = re.match(r"\s*launch__grid_size\s+([0-9]+)", line)
            if m:
                grid = int(m.group(1))
        if seen == 2:
            best = (tot, os.path.relpath(path, ROOT), regs, grid)
    if best is None:
        return None
    tot, path, regs, grid = best
    if launch is not None and (regs != launch["registers"] or grid != launch["grid"]):
        return (None, path, "stale capture: %s registers / grid %s, running kernel %d registers / grid %d" %
                (regs, grid, launch["registers"], launch["grid"]))
    return (tot, path, None)


# ---------------------------------------------------------------------------------------------
# the reference on the host CPU (test infrastructure under oracle/ is executed here, and only here)
# ---------------------------------------------------------------------------------------------
def run_reference_cpu(samples, warmup, steps):
    """Runs oracle/_ref/MC_ref on the HEU deck; returns (histories/s over `steps` generations, cores, sample text).
    Generation boundaries are timestamped from the reference's own per-cycle stdout lines (through a pty, so the
    lines arrive unbuffered)."""
    import pty
    from mc_old_b200 import decks
    ref_dir = os.path.join(ROOT, "oracle", "_ref")
    exe = os.path.join(ref_dir, "MC_ref")
    if not os.path.exists(exe):
        return None
    d = tempfile.mkdtemp(prefix="mcb_ref_")
    decks.write(d, decks.heu_sphere(samples=samples, active=steps, passive=warmup))
    master, slave = pty.openpty()
    p = subprocess.Popen([exe, d], cwd=ref_dir, stdout=slave, stderr=subprocess.STDOUT, close_fds=True)
    os.close(slave)
    stamps, buf = [], b""
    t_start = time.perf_counter()
    while True:
        try:
            chunk = os.read(master, 4096)
        except OSError:
            break
        if not chunk:
            break
        now = time.perf_counter()
        buf += chunk
        while b"\n" in buf:
            line, buf = buf.split(b"\n", 1)
            parts = line.strip().split()
            if len(parts) >= 2 and parts[0].isdigit():
                stamps.append(now)
            elif b"running the simulation" in line:
                t_start = now
    p.wait()
    os.close(master)
    if p.returncode != 0 or len(stamps) < warmup + steps:
        return None
    t0 = stamps[warmup - 1] if warmup > 0 else t_start
    dt = stamps[warmup + steps - 1] - t0
    return steps * samples / dt, 1, ("MC_ref (g++ -O3, 1 thread: the reference is single-threaded and not re-entrant), "
                                     "%d generations x %d histories after %d warm-up generations" % (steps, samples, warmup)), dt / steps


def reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    samples = args.ref_samples
    r = run_reference_cpu(samples, args.warmup, args.steps)
    if r is None:
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/MC_ref is not built (run __graft_entry__.build() where /root/reference exists)"}))
        return 0
    value, cores, sample, s_per_step = r
    line = {
        "impl": "reference", "metric": "histories_per_second", "value": value, "unit": "histories/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * s_per_step,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOAD, "histories_per_generation": samples, "note": "bounded sample of the same deck"},
        "reference_histories_per_generation": samples,
        "why_not_the_gpu_size": "the reference holds two banks of heap-allocated particles (~350 B per site): 1e7 histories per generation need ~7 GB and "
                                "~40 s per generation on one core (BASELINE.md section 2); histories/s is normalised by size and falls slowly with it "
                                "(see cpu_baseline.points of the GPU arm: 2e5 and 1e6 per generation)",
        "cpu_baseline": {"value": value, "unit": "histories/s", "cores": cores, "kind": "reference", "sample": sample},
        "e2e": {"value": value, "unit": "histories/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))
    return 0


# ---------------------------------------------------------------------------------------------
class ClockSampler:
    """SM clock and throttle reasons sampled DURING the timed region: NVML polled every few milliseconds from a
    thread (a generation is ~9 ms, `nvidia-smi -lms` cannot see it); nvidia-smi is the fallback."""
    BAD = {"hw_slowdown": 0x8, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20, "sw_power_cap": 0x4}

    def __init__(self, device):
        import threading
        self.sm, self.reasons, self.max_mhz = [], set(), None
        self.stop_flag = False
        self.smi = None
        try:
            import pynvml
            pynvml.nvmlInit()
            # NVML enumerates physical devices: honour CUDA_VISIBLE_DEVICES
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            idx = int(vis.split(",")[device]) if vis and all(x.strip().isdigit() for x in vis.split(",")) else device
            self.h = pynvml.nvmlDeviceGetHandleByIndex(idx)
            self.nv = pynvml
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
            self.t = threading.Thread(target=self._poll, daemon=True)
            self.t.start()
        except Exception:
            self.nv = None
            self.path = tempfile.mktemp(prefix="mcb_clocks_", suffix=".csv")
            self.f = open(self.path, "w")
            q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
                 "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
            try:
                self.smi = subprocess.Popen(["nvidia-smi", "-i", str(device), "--query-gpu=" + q, "--format=csv,noheader,nounits",
                                             "-lms", "20"], stdout=self.f, stderr=subprocess.DEVNULL)
            except Exception:
                self.smi = None

    def _poll(self):
        nv = self.nv
        while not self.stop_flag:
            try:
                self.sm.append(float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
                r = int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h))
                for name, bit in self.BAD.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.002)

    def stop(self):
        if self.nv is not None:
            self.stop_flag = True
            self.t.join(timeout=2)
        elif self.smi is not None:
            self.smi.terminate()
            try:
                self.smi.wait(timeout=5)
            except Exception:
                self.smi.kill()
            self.f.close()
            for line in open(self.path):
                c = [x.strip() for x in line.split(",")]
                try:
                    self.sm.append(float(c[0])); self.max_mhz = float(c[1])
                except (ValueError, IndexError):
                    continue
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), c[2:6]):
                    if v.lower().startswith("active"):
                        self.reasons.add(name)
            os.unlink(self.path)
        if not self.sm:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": 0}
        sm = sorted(self.sm)
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(sm)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=4)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--samples", type=float, default=1e7, help="histories per generation per GPU")
    ap.add_argument("--ref-samples", type=float, default=2e5, help="histories per generation of the CPU reference runs")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-xs", action="store_true", help="skip the xs_lookup microbench")
    ap.add_argument("--record-g-independence", action="store_true", help="write tests/golden/g_independence.json from this (1-GPU) run")
    args = ap.parse_args()
    args.ref_samples = int(args.ref_samples)
    if args.warmup < 3:
        args.warmup = 3
    if args.impl == "reference":
        return reference_arm(args)

    import numpy as np
    import torch
    import torch.distributed as dist
    import mc_old_b200 as mcb
    from mc_old_b200 import decks

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device: there is no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        dist.barrier()  # creates torch's NCCL communicator, so libnccl is loaded before ours binds it
    per_gpu = int(args.samples)
    n_sample = per_gpu * world
    total_cycles = args.warmup + 3 * args.steps + 8  # all passive but the warm-up: the HEU deck has no user tallies
    deck = mcb.Deck(xml=decks.heu_sphere(samples=n_sample, active=total_cycles, passive=args.warmup))
    stream = torch.cuda.current_stream()
    ctx = mcb.Context(deck, device=local_rank, rank=rank, world=world, stream=stream.cuda_stream)
    if world > 1:
        uid = [mcb.Context.unique_id() if rank == 0 else None]
        dist.broadcast_object_list(uid, src=0)
        ctx.comm_init(uid[0])

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- GPU-count independence check (untimed) ----
    gdeck = mcb.Deck(xml=decks.heu_sphere(samples=G_IND_SAMPLES, active=1, passive=G_IND_CYCLES - 1))
    gctx = mcb.Context(gdeck, device=local_rank, rank=rank, world=world, stream=stream.cuda_stream)
    if world > 1:
        uid = [mcb.Context.unique_id() if rank == 0 else None]
        dist.broadcast_object_list(uid, src=0)
        gctx.comm_init(uid[0])
    g_got = []
    for _ in range(G_IND_CYCLES):
        r = gctx.run_cycle()
        g_got.append([float(r.k_cycle).hex(), int(r.n_sites), int(r.n_tracks)])
    gctx.close()
    g_independent, g_expected = None, None
    if args.record_g_independence and world == 1:
        with open(G_IND_FILE, "w") as f:
            json.dump({"samples": G_IND_SAMPLES, "cycles": G_IND_CYCLES, "deck": "HEU_sphere_criticality", "recorded_with_gpus": 1,
                       "k_cycle_hex_n_sites_n_tracks": g_got}, f, indent=1)
    if os.path.exists(G_IND_FILE):
        g_expected = json.load(open(G_IND_FILE))["k_cycle_hex_n_sites_n_tracks"]
        g_independent = g_got == g_expected

    for _ in range(args.warmup):
        ctx.run_cycle()

    # ---- timed region: K generations, device time on the launch stream ----
    sampler = ClockSampler(local_rank) if rank == 0 else None
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    res = []
    barrier()
    e0.record(stream)
    for _ in range(args.steps):
        res.append(ctx.run_cycle())
    e1.record(stream)
    barrier()
    ms = max_over_ranks(e0.elapsed_time(e1))
    clocks = sampler.stop() if sampler else None
    hist = sum(r.n_histories for r in res)          # global counts (summed over the ranks inside the library)
    coll = sum(r.n_collisions for r in res)
    tracks = sum(r.n_tracks for r in res)
    lookups = sum(r.n_lookups for r in res)
    launches = sum(r.n_kernel_launches for r in res)
    value = hist / (ms * 1e-3)

    # ---- e2e: the same step with the source bank crossing the host boundary both ways ----
    e2e = None
    if not args.no_e2e:
        cap = (4 if world == 1 else 2) * n_sample + 4096 * world
        bufs = [(torch.empty((cap, 8), dtype=torch.float64, pin_memory=True).numpy(), torch.empty((cap,), dtype=torch.int32, pin_memory=True).numpy())
                for _ in range(2)]
        s, c = ctx.source_bank(cap, bufs[0][0], bufs[0][1])
        n_bank = s.shape[0]
        # one untimed step: the copy stream, staging and sort buffers of the host-bank path are created on first use
        r, s, c = ctx.run_cycle_host(bufs[0][0][:n_bank], bufs[0][1][:n_bank], bufs[1][0], bufs[1][1])
        n_bank = s.shape[0]
        bufs.reverse()
        h2d = d2h = 0
        barrier()
        t0 = time.perf_counter()
        for i in range(args.steps):
            src, dst = bufs[i & 1], bufs[(i + 1) & 1]
            h2d += mcb.shard_range(n_bank, rank, world)[1] * 68   # with several GPUs every rank moves only its slice
            r, s, c = ctx.run_cycle_host(src[0][:n_bank], src[1][:n_bank], dst[0], dst[1])
            n_bank = s.shape[0]
            d2h += mcb.shard_range(n_bank, rank, world)[1] * 68 + 176
        barrier()
        dt = max_over_ranks(time.perf_counter() - t0)
        e2e = {"value": args.steps * n_sample / dt, "unit": "histories/s", "h2d_bytes_per_step": h2d // args.steps,
               "d2h_bytes_per_step": d2h // args.steps, "ms_per_step": 1e3 * dt / args.steps,
               "what": "per step and per rank: mcb_run_cycle_host = source bank from pinned host memory -> HBM, the generation, the new "
                       "bank HBM -> pinned host memory.  One GPU: the whole bank, upload pipelined with the walk.  Several GPUs: the "
                       "host bank is one array of which every rank moves its 1/W slice each way (bytes are per rank); the slices "
                       "are read in place over NVLink"}

    # ---- per-stage pass: CUDA events around every stage launch (rank-local), for the roofline of the dominant kernel ----
    ctx_launch = ctx.walk_launch_info(False)
    ctx.reset_stage_times()
    ctx.set_stage_timing(True)
    for _ in range(2):
        ctx.run_cycle()
    ctx.set_stage_timing(False)
    st = ctx.stage_times()
    nn = 2
    stage_bytes = {"lookup": st["units_lookup"] * BYTES_LOOKUP(nn)}
    # units of the other stages on this rank over the two profiled generations (rank-local shares of the global counts)
    frac = 1.0 / world
    stages = {k: {"ms": st["ms_" + k], "launches": st["n_" + k]} for k in ("source", "lookup", "flight", "cross", "collide", "step", "finish", "closeout", "bank")}
    dominant = max(("lookup", "flight", "collide", "cross", "step"), key=lambda k: stages[k]["ms"])
    peak, peak_kind = peaks()
    roofline = None
    if stages[dominant]["ms"] > 0:
        # unit counts of the profiled generations: scale the timed-region averages per generation
        per_gen = {"lookup": st["units_lookup"] / 2.0, "flight": tracks / args.steps * frac, "collide": coll / args.steps * frac,
                   "cross": (tracks - coll) / args.steps * frac, "step": tracks / args.steps * frac}
        per_unit = {"lookup": BYTES_LOOKUP(nn), "flight": BYTES_FLIGHT, "collide": BYTES_COLLIDE(nn), "cross": BYTES_CROSS,
                    "step": BYTES_STEP(nn)}
        total_bytes = 2.0 * per_gen[dominant] * per_unit[dominant]
        achieved = total_bytes / (stages[dominant]["ms"] * 1e-3) / 1e9
        kname = "k_" + ("xs_stage" if dominant == "lookup" else ("walk" if dominant == "step" else dominant))
        launch = ctx_launch if dominant == "step" else None
        tr = ncu_traffic(kname, launch) if int(args.samples) == 10000000 else None
        roofline = {"bound": "hbm", "kernel": "k_" + ("xs_stage" if dominant == "lookup" else ("walk" if dominant == "step" else dominant)), "unit_name": "lookup" if dominant == "lookup" else ("collision" if dominant == "collide" else "track"), "achieved": achieved,
                    "peak": peak, "peak_kind": peak_kind, "unit": "GB/s", "frac": achieved / peak,
                    "traffic": tr[0] if tr else None, "traffic_source": tr[1] if tr else None, "traffic_refused": tr[2] if tr else None,
                    "launch": launch,
                    "algorithmic_bytes_per_launch": total_bytes / max(stages[dominant]["launches"], 1),
                    "bytes_per_unit": per_unit[dominant], "units_per_launch": 2.0 * per_gen[dominant] / max(stages[dominant]["launches"], 1),
                    "avg_launch_ms": stages[dominant]["ms"] / max(stages[dominant]["launches"], 1),
                    "note": "bytes_per_unit is SURVEY 8(d)'s fixed per-track figure (record in+out 224, lookup 72+100*Nn, sites 20); the walk kernel keeps the record in registers and the tables in L2, so its DRAM traffic is below it (DESIGN.md 4)"}
    # ---- xs_lookup microbench (SURVEY 8d input 1): macroscopic cross sections of the HEU material at 2^26 energies,
    # log-uniform on [1e-5, 2e7] eV and Watt-shaped (what the transport loop asks for); device-resident in and out ----
    xs_micro = None
    if rank == 0 and not args.no_xs:
        n_e = 1 << 26
        g = torch.Generator(device="cuda"); g.manual_seed(12345)
        out5 = torch.empty((n_e, 5), dtype=torch.float64, device="cuda")
        xs_micro = {}
        for label in ("log_uniform", "watt_spectrum"):
            u = torch.rand(n_e, dtype=torch.float64, device="cuda", generator=g)
            if label == "log_uniform":
                E = torch.exp(np.log(1e-5) + u * (np.log(2e7) - np.log(1e-5)))
            else:  # Maxwellian-like fission spectrum, T = 1.3 MeV: E = -T (ln u1 + ln u2 cos^2(pi u3 / 2))
                u2 = torch.rand(n_e, dtype=torch.float64, device="cuda", generator=g)
                u3 = torch.rand(n_e, dtype=torch.float64, device="cuda", generator=g)
                E = -1.3e6 * (torch.log(u) + torch.log(u2) * torch.cos(0.5 * np.pi * u3) ** 2)
            torch.cuda.synchronize()
            best = min(ctx.xs_lookup_device(0, E.data_ptr(), n_e, out5.data_ptr()) for _ in range(5))
            bytes_alg = n_e * BYTES_LOOKUP(nn)
            xs_micro[label] = {"lookups_per_second": n_e / (best * 1e-3), "ms": best, "algorithmic_GBps": bytes_alg / (best * 1e-3) / 1e9,
                               "io_GBps": n_e * 48 / (best * 1e-3) / 1e9, "io_frac_of_hbm_peak": n_e * 48 / (best * 1e-3) / 1e9 / peak}
        # the roofline fraction of this kernel is its real I/O (8 B in + 40 B out per lookup) against the HBM peak; the
        # SURVEY 8(d) algorithmic figure counts table bytes that L2 / L1 serve and is not a roofline (it can exceed the peak)
        xs_micro["bound"] = ("L1 gather pipe: l1tex__data_pipe_lsu_wavefronts 89 % of peak (profiles/r2f_k_xs_lookup_ncu_full_summary.txt); "
                             "every lane gathers 16-byte pieces of its own table rows from a different line")
        del out5
    ctx.close()

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        r = run_reference_cpu(args.ref_samples, 1, 4)
        if r is not None:
            cpu = {"value": r[0], "unit": "histories/s", "cores": r[1], "kind": "reference", "sample": r[2],
                   "reference_histories_per_generation": args.ref_samples,
                   "why_not_the_gpu_size": "the reference needs ~7 GB and ~40 s per generation at 1e7 histories per generation (BASELINE.md section 2)",
                   "points": [{"histories_per_generation": args.ref_samples, "histories_per_second": r[0]}]}
            r2 = run_reference_cpu(1000000, 1, 2)  # a second size, so that the size dependence is data, not a claim
            if r2 is not None:
                cpu["points"].append({"histories_per_generation": 1000000, "histories_per_second": r2[0]})

    if rank == 0:
        line = {
            "metric": "histories_per_second", "value": value, "unit": "histories/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": WORKLOAD, "histories_per_generation": n_sample, "histories_per_gpu": per_gpu,
                       "parallelism": "histories sharded over %d GPU(s); per generation one NCCL all-gather of the close-out sums, fission bank read in place over NVLink" % world,
                       "l2": "inputs larger than L2: per GPU and generation the source bank (%.2f GB), particle bank (%.2f GB) and site requests (%.2f GB) stream through HBM" % (per_gpu * 64 / 1e9, per_gpu * 100 / 1e9, per_gpu * 64 / 1e9)},
            "collisions_per_second": coll / (ms * 1e-3), "tracks_per_second": tracks / (ms * 1e-3),
            "xs_lookups_per_second": lookups / (ms * 1e-3),
            "xs_lookup_algorithmic_GBps": lookups * BYTES_LOOKUP(2) / world / (ms * 1e-3) / 1e9,
            "xs_lookup_microbench": xs_micro,
            "k_cycle_last": res[-1].k_cycle, "event_loop_iterations_per_step": sum(r.n_iterations for r in res) / args.steps,
            "g_independent": g_independent, "g_independence": {"got": g_got, "expected": g_expected, "golden": os.path.relpath(G_IND_FILE, ROOT)},
            "gpu_launches": int(launches), "clocks": clocks, "e2e": e2e, "roofline": roofline, "stages_ms_2_generations": stages,
            "cpu_baseline": cpu,
        }
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if g_independent is False:
        sys.stderr.write("bench.py: results depend on the GPU count: got %r, expected %r\n" % (g_got, g_expected))
        return 3
    return 0


if __name__ == "__main__":
    sys.exit(main())
