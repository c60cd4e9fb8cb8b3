#!/usr/bin/env python
"""Pins oracle/replay.py::Memory to the reference's utils/memory.py::Memory (run in the build container:
needs /root/reference).  Writes tests/golden/replay_memory.npz: a stream of experiences pushed through a
small ring and the windows the reference samples with a seeded numpy generator."""
import os
import sys

import numpy as np

REF = os.environ.get("DIRAL_REFERENCE", "/root/reference")
sys.path.insert(0, REF)
from utils.memory import Memory  # noqa: E402  (the reference class)

N, S, CAP, T, BATCH, STEP = 5, 7, 12, 31, 4, 3
rs = np.random.RandomState(11)
states = rs.rand(T + 1, N, S)
actions = rs.randint(0, 4, size=(T, N))
rewards = rs.randn(T, N)
mem = Memory(max_size=CAP)
samples, idxs, lens = [], [], []
for t in range(T):
    mem.add((states[t], actions[t], rewards[t], states[t + 1]))
    if t >= BATCH + STEP and t % 3 == 0:
        np.random.seed(100 + t)
        # the reference draws with the global numpy generator (utils/memory.py:184)
        batch = mem.sample(BATCH, STEP)
        np.random.seed(100 + t)
        idx = np.random.choice(np.arange(len(mem.buffer) - STEP), size=BATCH, replace=False)
        samples.append(np.array([[np.concatenate([e[0].ravel(), e[1].ravel(), e[2].ravel(), e[3].ravel()]) for e in w]
                                 for w in batch]))
        idxs.append(idx); lens.append(t)
out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "replay_memory.npz")
np.savez_compressed(out, states=states, actions=actions, rewards=rewards, samples=np.stack(samples), idx=np.stack(idxs),
                    at=np.array(lens), shape=np.array([N, S, CAP, T, BATCH, STEP]))
print(out, len(samples))
