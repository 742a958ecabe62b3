#!/usr/bin/env python3
"""xs_lookup microbench alone (bench.py's, same inputs): tools/xs_rate.py  [MCB200_LIB=variant.so]  -> one line per spectrum"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mc_old_b200 as mcb
from mc_old_b200 import decks

deck = mcb.Deck(xml=decks.heu_sphere(samples=1000, active=1, passive=1))
ctx = mcb.Context(deck, device=0)
n_e = 1 << 26
g = torch.Generator(device="cuda"); g.manual_seed(12345)
out5 = torch.empty((n_e, 5), dtype=torch.float64, device="cuda")
res = []
for label in ("log_uniform", "watt_spectrum"):
    u = torch.rand(n_e, dtype=torch.float64, device="cuda", generator=g)
    if label == "log_uniform":
        E = torch.exp(np.log(1e-5) + u * (np.log(2e7) - np.log(1e-5)))
    else:
        u2 = torch.rand(n_e, dtype=torch.float64, device="cuda", generator=g)
        u3 = torch.rand(n_e, dtype=torch.float64, device="cuda", generator=g)
        E = -1.3e6 * (torch.log(u) + torch.log(u2) * torch.cos(0.5 * np.pi * u3) ** 2)
    torch.cuda.synchronize()
    best = min(ctx.xs_lookup_device(0, E.data_ptr(), n_e, out5.data_ptr()) for _ in range(5))
    res.append("%s %.3f ms %.2f G lookups/s (sum %.17g)" % (label, best, n_e / best / 1e6, float(out5.sum())))
ctx.close()
print(os.path.basename(os.environ.get("MCB200_LIB", "libmcb200.so")), " | ".join(res))
