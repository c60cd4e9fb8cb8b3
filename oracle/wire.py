"""TEST INFRASTRUCTURE, not product code: CPU restatement of the RealNeS environment's view-based positional
distribution (envs/realness_env.py:52-118, :193-207) and of the semi-persistent-scheduling baseline
(algorithms/v2x_sps.py:4-104), in the reference's own loop order and float64 arithmetic.

Parity status: pinned.  tests/golden/make_golden_wire.py executes the unmodified reference methods (the ZMQ /
protobuf imports stubbed, the `random` module of v2x_sps replaced by a scripted one) and records
tests/golden/wire_vpd.npz and tests/golden/sps.npz; tests/test_oracle_golden.py replays them bit for bit.
"""
import math

import numpy as np


def _dist(rx, tx):
    """RealnessEnv.dist (realness_env.py:193-207)."""
    (x1, y1), (x2, y2) = rx, tx
    d = math.sqrt((x2 - x1) ** 2 + (y2 - y1) ** 2)
    return d, (1 if x1 - x2 > 0.0 else -1)


def _samples(tx_id, xpos, ypos, last_updated, age_limit):
    out = []
    for rx_id in range(len(xpos)):                                        # realness_env.py:62-72 / :97-107
        if tx_id == rx_id:
            continue
        if last_updated[rx_id] > age_limit:
            continue
        d, sign = _dist((float(xpos[rx_id]), float(ypos[rx_id])), (float(xpos[tx_id]), float(ypos[tx_id])))
        out.append(d * sign)
    return out


def neighbor_dist2(tx_id, xpos, ypos, last_updated, state_bins, state_range, age_limit=20):
    """RealnessEnv.get_neighbor_dist2 (realness_env.py:87-118)."""
    s = _samples(tx_id, xpos, ypos, last_updated, age_limit)
    if len(s) > 0:
        h = np.histogram(sorted(s), state_bins, range=(-state_range, state_range))[0]
        return h / float(len(s))
    return np.zeros((state_bins,), dtype=int)


def neighbor_dist(tx_id, xpos, ypos, last_updated, state_bins, age_limit=20):
    """RealnessEnv.get_neighbor_dist (realness_env.py:52-85)."""
    s = _samples(tx_id, xpos, ypos, last_updated, age_limit)
    if len(s) > 0:
        bins = np.linspace(-1, 1, state_bins + 1)
        srt = sorted(s)
        norm = np.linalg.norm(s, np.inf)
        normed = srt / norm
        return np.histogram(normed, bins, weights=normed)[0]
    return np.zeros((state_bins,), dtype=int)


class SpsBank:
    """A bank of SemiPersistentScheduling agents (v2x_sps.py:4-104) driven by explicit draws:
    draws[a] = (new reselection counter, keep-uniform, choice index)."""

    def __init__(self, prev_action, reselection_counter, rssi_threshold, inc_db=3, prob_keep=0.8):
        self.prev = np.array(prev_action, dtype=np.int64)
        self.counter = np.array(reselection_counter, dtype=np.int64)
        self.thr, self.inc, self.keep = rssi_threshold, inc_db, prob_keep

    def _choose(self, a, window, index):
        """choose_new_resource (v2x_sps.py:24-74); returns None where the reference would raise / spin."""
        min_sa = len(window) / 5                                          # :40 (true division on Python 3)
        tmp, s_a, guard = self.thr, {}, 0
        while len(s_a) < min_sa:                                          # :42-51
            s_a = {}
            for sub in range(len(window)):
                if self.prev[a] == sub:
                    continue
                if window[sub] < tmp:
                    s_a[sub] = window[sub]
            tmp += self.inc
            guard += 1
            if guard > 4097:
                return None
        srt = sorted(s_a.items(), key=lambda kv: kv[1])                   # :54
        min_len = min(min_sa, len(s_a))
        s_b = []
        for k, _ in srt:                                                  # :56-59
            s_b.append(k)
            if len(s_b) >= min_len:
                break
        if not s_b:
            return None
        return s_b[int(index) % len(s_b)]                                 # :60 with a scripted random.choice

    def step(self, windows, draws):
        acts = np.empty(len(self.prev), dtype=np.int64)
        flags = np.zeros(len(self.prev), dtype=np.int64)
        for a in range(len(self.prev)):                                   # v2x_sps.py:76-104
            if self.counter[a] != 0:
                acts[a] = self.prev[a]
                self.counter[a] -= 1
                continue
            self.counter[a] = int(draws[a][0])
            if draws[a][1] < self.keep:
                acts[a] = self.prev[a]
            else:
                c = self._choose(a, windows[a], draws[a][2])
                if c is None:
                    flags[a] = 1
                    acts[a] = self.prev[a]
                else:
                    acts[a] = c
                    self.prev[a] = c
        return acts, flags
