"""Worker of tests/test_gpu_transport.py::test_multi_gpu_matches_single_gpu: one rank per GPU under torchrun.
torch.distributed carries the NCCL unique id; the library does its own NCCL collectives."""
import argparse
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mc_old_b200 as mcb
from mc_old_b200 import decks

ap = argparse.ArgumentParser()
ap.add_argument("--samples", type=int, default=40000)
ap.add_argument("--cycles", type=int, default=3)
a = ap.parse_args()
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
dist.barrier()
deck = mcb.Deck(xml=decks.heu_sphere(samples=a.samples, active=a.cycles - 1, passive=1, entropy=True, estimators=True))
ctx = mcb.Context(deck, device=local, rank=rank, world=world)
uid = [mcb.Context.unique_id() if rank == 0 else None]
dist.broadcast_object_list(uid, src=0)
ctx.comm_init(uid[0])
rs = [ctx.run_cycle() for _ in range(a.cycles)]
sites, _ = ctx.source_bank(int(rs[-1].n_sites))
mean, uncer = ctx.tallies()
out = {"k_cycle_hex": [r.k_cycle.hex() for r in rs], "n_sites": [int(r.n_sites) for r in rs],
       "H_hex": [r.H.hex() for r in rs], "n_tracks": [int(r.n_tracks) for r in rs],
       "bank_energy_sum_hex": float(np.sum(sites[:, 6])).hex(),
       "tally_mean": [float(x) for x in mean], "tally_uncer": [float(x) for x in uncer]}
# host-bank cycles (mcb_run_cycle_host): every rank reads its slice of the global host bank and writes its slice of the new one
cells = ctx.source_bank(int(rs[-1].n_sites))[1]
hk = []
for _ in range(2):
    out_s = np.zeros((4 * a.samples, 8)); out_c = np.zeros(4 * a.samples, dtype=np.int32)
    r, s_new, c_new = ctx.run_cycle_host(sites, cells, out_s, out_c)
    parts = [None] * world
    dist.all_gather_object(parts, (s_new.copy(), c_new.copy()))
    if os.environ.get("MCB_NO_P2P"):   # no peer mappings: every rank moved the whole bank
        sites, cells = parts[0]
    else:                               # the slices are disjoint, the rest of every rank's array is zero
        sites = sum(p[0] for p in parts)
        cells = sum(p[1] for p in parts).astype(np.int32)
    hk.append([r.k_cycle.hex(), int(r.n_sites), int(r.n_tracks)])
out["host_cycles"] = hk
out["host_bank_sum_hex"] = [float(np.sum(sites[:, i])).hex() for i in range(8)]
gathered = [None] * world
dist.all_gather_object(gathered, out)
assert all(g == gathered[0] for g in gathered), "ranks disagree on the global results"
ctx.close()
if rank == 0:
    print(json.dumps(out), flush=True)
dist.barrier()
dist.destroy_process_group()
