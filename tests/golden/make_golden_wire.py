#!/usr/bin/env python
"""Pins oracle/wire.py to the unmodified reference (run in the build container: needs /root/reference).

* wire_vpd.npz: RealnessEnv.get_neighbor_dist / get_neighbor_dist2 (envs/realness_env.py:52-118) called on
  random neighbour tables in the dict-of-dict form RealNeSZmqBridge.get_observation_syn_dist builds
  (envs/realness_bridge.py:168-191), positions rounded to float32 as on the wire.  ZMQ and the generated
  protobuf module are stubbed: neither is touched by these two methods.
* sps.npz: SemiPersistentScheduling (algorithms/v2x_sps.py) stepped with its `random` module replaced by a
  scripted one, so the draws are part of the fixture.
"""
import os
import sys
import types
from collections import defaultdict

import numpy as np

REF = os.environ.get("DIRAL_REFERENCE", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))
for name in ("zmq", "ma_messages_pb2", "matplotlib", "matplotlib.pyplot"):
    sys.modules.setdefault(name, types.ModuleType(name))
sys.path.insert(0, os.path.join(REF, "envs"))
sys.path.insert(0, os.path.join(REF, "algorithms"))
from realness_env import RealnessEnv  # noqa: E402  (the reference class; never instantiated: __init__ opens sockets)
import v2x_sps  # noqa: E402


def ref_self(bins, rng):
    s = types.SimpleNamespace(state_bins=bins, state_range=rng)
    s.dist = lambda p1, p2: RealnessEnv.dist(s, p1, p2)
    return s


def make_vpd():
    rs = np.random.RandomState(5)
    cases = []
    for (M, N, bins, rng, span) in [(40, 6, 10, 250.0, 1200.0), (40, 12, 10, 250.0, 600.0), (30, 40, 20, 500.0, 2500.0),
                                     (10, 33, 7, 100.0, 150.0), (6, 1, 10, 250.0, 10.0)]:
        x = rs.uniform(0, span, size=(M, N)).astype(np.float32)
        y = rs.choice(np.array([0.0, 3.5, 7.0, 10.5], dtype=np.float32), size=(M, N))
        seq = rs.randint(0, 500, size=(M, N)).astype(np.int32)
        lu = rs.randint(0, 41, size=(M, N)).astype(np.int32)
        lu[rs.rand(M) < 0.1] = 30                     # whole tables stale: the all-zero branch
        obs = rs.randint(0, N, size=M).astype(np.int32)
        x[0, :] = x[0, 0]; y[0, :] = y[0, 0]          # everybody on one spot: zero distances (type 2 only)
        if N > 2:
            x[1, 1] = np.float32(x[1, obs[1]] + rng) if obs[1] != 1 else x[1, 1]   # a sample exactly on the last edge
            y[1, 1] = y[1, obs[1]]
            lu[1, 1] = 0
        o1 = np.zeros((M, bins)); o2 = np.zeros((M, bins))
        for m in range(M):
            tab = defaultdict(dict)
            for u in range(N):
                tab[u]["xpos"] = float(x[m, u]); tab[u]["ypos"] = float(y[m, u])
                tab[u]["seq_number"] = int(seq[m, u]); tab[u]["last_updated"] = int(lu[m, u])
            s = ref_self(bins, rng)
            o2[m] = RealnessEnv.get_neighbor_dist2(s, int(obs[m]), tab)
            if m > 0:
                with np.errstate(all="ignore"):
                    o1[m] = RealnessEnv.get_neighbor_dist(s, int(obs[m]), tab)
        cases.append(dict(x=x, y=y, seq=seq, lu=lu, obs=obs, bins=bins, rng=rng, o1=o1, o2=o2))
    out = {}
    for i, c in enumerate(cases):
        for k, v in c.items():
            out["c%d_%s" % (i, k)] = v
    out["ncases"] = len(cases)
    np.savez_compressed(os.path.join(HERE, "wire_vpd.npz"), **out)
    print("wire_vpd.npz", len(cases), "cases")


class ScriptedRandom:
    """Stands in for the `random` module inside v2x_sps: every call consumes the scripted triple of the
    current (agent, step): randint -> d0 (or the init values), random -> u1, choice -> seq[d2 % len]."""
    def __init__(self):
        self.cur = None
        self.init = None

    def randint(self, a, b):
        if self.init is not None:
            return self.init.pop(0)
        return int(self.cur[0])

    def random(self):
        return float(self.cur[1])

    def choice(self, seq):
        return seq[int(self.cur[2]) % len(seq)]


def make_sps():
    rs = np.random.RandomState(9)
    fake = ScriptedRandom()
    v2x_sps.random = fake
    out = {}
    for ci, (A, Wn, T, thr) in enumerate([(24, 20, 60, -97.0), (16, 12, 60, -97.0), (8, 5, 40, -110.0)]):
        tx0 = rs.randint(0, Wn + 1, size=A); c0 = rs.randint(5, 16, size=A)
        agents = []
        for a in range(A):
            fake.init = [int(tx0[a]), int(c0[a])]
            agents.append(v2x_sps.SemiPersistentScheduling(a, Wn, thr))
            fake.init = None
        windows = rs.uniform(-125.0, -70.0, size=(T, A, Wn))
        windows[rs.rand(T, A, Wn) < 0.2] = -117.0          # ties in the RSSI ordering
        draws = np.stack([rs.randint(5, 17, size=(T, A)).astype(np.float64), rs.rand(T, A),
                          rs.randint(0, 1 << 20, size=(T, A)).astype(np.float64)], axis=-1)
        draws[..., 0][rs.rand(T, A) < 0.5] = 0.0           # short counters: more reselections in T steps
        acts = np.zeros((T, A), dtype=np.int64); prev = np.zeros((T, A), dtype=np.int64); cnt = np.zeros((T, A), dtype=np.int64)
        for t in range(T):
            for a in range(A):
                fake.cur = draws[t, a]
                acts[t, a] = agents[a].step(list(windows[t, a]))
                prev[t, a] = agents[a].prev_action; cnt[t, a] = agents[a].reselection_counter
        out.update({"c%d_tx0" % ci: tx0, "c%d_c0" % ci: c0, "c%d_windows" % ci: windows, "c%d_draws" % ci: draws,
                    "c%d_acts" % ci: acts, "c%d_prev" % ci: prev, "c%d_cnt" % ci: cnt, "c%d_thr" % ci: thr})
    out["ncases"] = 3
    np.savez_compressed(os.path.join(HERE, "sps.npz"), **out)
    print("sps.npz")


if __name__ == "__main__":
    make_vpd()
    make_sps()
