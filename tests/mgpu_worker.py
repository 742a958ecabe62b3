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
gathered = [None] * world
dist.all_gather_object(gathered, out)
assert all(g == gathered[0] for g in gathered), "ranks disagree on the global results"
ctx.close()
if rank == 0:
    print(json.dumps(out), flush=True)
dist.barrier()
dist.destroy_process_group()
