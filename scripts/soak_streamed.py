#!/usr/bin/env python
"""Soak test of the streamed host records (device-raised chunk flags, one system-scope fence per chunk): the same
environments stepped through the chunked copy-engine format and through diral_step_host_begin / _wait, every row of
every slot compared.  python scripts/soak_streamed.py [slots]"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from bench import ENV_KW, N_UE  # noqa: E402
from diral_b200 import TestEnv  # noqa: E402

slots = int(sys.argv[1]) if len(sys.argv) > 1 else 5000
G, E = 4, 1024
ref = [TestEnv(num_envs=E, device="cuda:0", seed=7 + g, host_format="compact", host_threads=4, **ENV_KW) for g in range(G)]
pipe = [TestEnv(num_envs=E, device="cuda:0", seed=7 + g, **ENV_KW) for g in range(G)]
for e in pipe:
    e.set_host_format("compact_stream", 8, shared_pool=True)
    e.host_stream = torch.cuda.Stream()
S = ref[0].S
mk = lambda: (torch.empty((E, N_UE, S)).pin_memory(), torch.empty((E, N_UE)).pin_memory())
rb, pb = [mk() for _ in range(G)], [mk() for _ in range(G)]
acts = [[ref[g].sample(t).cpu().pin_memory() for t in range(16)] for g in range(G)]
torch.cuda.synchronize()
t0 = time.time()
for g in range(G):
    pipe[g].step_host_begin(acts[g][0], *pb[g])
bad = 0
for t in range(slots):
    for g in range(G):
        ref[g].step_host(acts[g][t % 16], *rb[g])
        pipe[g].step_host_wait()
        if not (torch.equal(pb[g][0], rb[g][0]) and torch.equal(pb[g][1], rb[g][1])):
            bad += 1
            print("MISMATCH slot", t, "group", g, flush=True)
        pb[g][0].fill_(-1.0)
        if t + 1 < slots:
            pipe[g].step_host_begin(acts[g][(t + 1) % 16], *pb[g])
print("soak: %d slots x %d groups of %d envs, %d mismatches, %.0f s" % (slots, G, E, bad, time.time() - t0))
sys.exit(1 if bad else 0)
