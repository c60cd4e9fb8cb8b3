"""ctypes binding of oracle/_build/libdiral_oracle.so -- TEST INFRASTRUCTURE, not product code.

``COracle(**EnvironmentTest, num_envs=E)`` mirrors the reference env surface
(envs/test_env.py:116-595) over a batch of E independent envs in float64.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "libdiral_oracle.so")

MODES = {"my_step": 0, "my_step_design": 1, "my_step_ch": 2}


class OrcCfg(C.Structure):
    _fields_ = [("N", C.c_int32), ("R", C.c_int32), ("B", C.c_int32),
                ("L", C.c_double), ("C", C.c_double), ("W", C.c_double),
                ("reward_design", C.c_int32), ("state_type", C.c_int32), ("toy", C.c_int32),
                ("mobility", C.c_int32), ("mobility_vary", C.c_int32), ("design_topology", C.c_int32),
                ("add_action", C.c_int32), ("action_binary", C.c_int32), ("add_channel_obs", C.c_int32),
                ("add_reward", C.c_int32), ("add_index", C.c_int32), ("add_velocity", C.c_int32),
                ("add_position", C.c_int32), ("add_positional_dist", C.c_int32), ("add_piggy", C.c_int32),
                ("pos_dist_type", C.c_int32), ("fingerprint", C.c_int32),
                ("age_threshold", C.c_int32), ("sentinel", C.c_double)]


class OrcBatch(C.Structure):
    _fields_ = [("E", C.c_int64),
                ("pos_x", C.c_void_p), ("pos_y", C.c_void_p), ("vel", C.c_void_p),
                ("tab_x", C.c_void_p), ("tab_y", C.c_void_p),
                ("tab_seq", C.c_void_p), ("tab_lu", C.c_void_p), ("lat", C.c_void_p),
                ("trace", C.c_void_p), ("trace_len", C.c_int64)]


def build(force: bool = False) -> str:
    """Compile the C oracle (gcc) if the shared object is missing or stale."""
    src = os.path.join(_HERE, "diral_oracle.c")
    hdr = os.path.join(_HERE, "diral_oracle.h")
    stale = (not os.path.exists(_LIB_PATH)
             or os.path.getmtime(_LIB_PATH) < max(os.path.getmtime(src), os.path.getmtime(hdr)))
    if force or stale:
        subprocess.check_call(["make", "-C", _HERE, "-s"] + (["-B"] if force else []))
    return _LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(build())
        _lib.orc_state_space.restype = C.c_int
    return _lib


def cfg_from_kwargs(**kw) -> OrcCfg:
    """Same defaults as TestEnv.__init__ (test_env.py:12-48)."""
    st = kw["State"]
    c = OrcCfg()
    c.N = int(kw.get("num_users", 3)); c.R = int(kw.get("num_channels", 3)); c.B = int(st["num_bins"])
    c.L = float(kw.get("highway_length", 200)); c.C = float(kw.get("communication_range", 1))
    c.W = float(kw.get("bin_range", 500))
    c.reward_design = int(kw.get("reward_design", 1)); c.state_type = int(st["type"])
    c.toy = int(bool(kw.get("congestion_test", False)))
    c.mobility = int(bool(kw.get("mobility", False))); c.mobility_vary = int(bool(kw.get("mobility_vary", False)))
    c.design_topology = int(bool(kw.get("enable_design_topology", False)))
    c.add_action = int(bool(st["add_action"])); c.action_binary = int(st["action_index"] == "binary")
    c.add_channel_obs = int(bool(st["add_channel_obs"])); c.add_reward = int(bool(st["add_reward"]))
    c.add_index = int(bool(st["add_index"])); c.add_velocity = int(bool(st["add_velocity"]))
    c.add_position = int(bool(st["add_position"])); c.add_positional_dist = int(bool(st["add_positional_dist"]))
    c.add_piggy = int(bool(st["add_positional_dist_piggy"])); c.pos_dist_type = int(st["add_positional_dist_type"])
    c.fingerprint = int(bool(kw.get("enable_fingerprint", False)))
    c.age_threshold = 20; c.sentinel = 100000.0
    return c


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


class COracle:
    """Batch of E float64 reference-semantics envs on the CPU."""

    def __init__(self, num_envs=1, threads=1, **kwargs):
        self.cfg = cfg_from_kwargs(**kwargs)
        self.E = int(num_envs)
        n = self.cfg.N
        self.N, self.R, self.B = n, self.cfg.R, self.cfg.B
        self.S = lib().orc_state_space(C.byref(self.cfg))
        E = self.E
        self.pos_x = np.zeros((E, n)); self.pos_y = np.zeros((E, n)); self.vel = np.zeros((E, n))
        self.tab_x = np.zeros((E, n, n)); self.tab_y = np.zeros((E, n, n))
        self.tab_seq = np.zeros((E, n, n), np.int32); self.tab_lu = np.zeros((E, n, n), np.int32)
        self.lat = np.full((E, n, n), -1, np.int32)
        self.trace = None
        self.threads = threads
        self._batch = OrcBatch()
        self._sync()

    def _sync(self):
        b = self._batch
        b.E = self.E
        for name in ("pos_x", "pos_y", "vel", "tab_x", "tab_y", "tab_seq", "tab_lu", "lat"):
            setattr(b, name, _p(getattr(self, name)))
        b.trace = _p(self.trace)
        b.trace_len = 0 if self.trace is None else self.trace.shape[0]

    def load_trace(self, trace):
        """Network.load_x_positions (network.py:171-178) with an in-memory [T,N] array."""
        self.trace = np.ascontiguousarray(trace, dtype=np.float64)
        self._sync()

    def reset(self, x0, y0, v0):
        x0 = np.ascontiguousarray(np.broadcast_to(x0, (self.E, self.N)), dtype=np.float64)
        y0 = np.ascontiguousarray(np.broadcast_to(y0, (self.E, self.N)), dtype=np.float64)
        v0 = np.ascontiguousarray(np.broadcast_to(v0, (self.E, self.N)), dtype=np.float64)
        lib().orc_reset(C.byref(self.cfg), C.byref(self._batch), _p(x0), _p(y0), _p(v0))

    def reset_philox(self, seed, env0=0):
        x0 = np.zeros((self.E, self.N)); y0 = np.zeros_like(x0); v0 = np.zeros_like(x0)
        lib().orc_philox_topology(C.byref(self.cfg), C.c_uint64(seed), C.c_int64(env0), C.c_int64(self.E),
                                  _p(x0), _p(y0), _p(v0))
        self.reset(x0, y0, v0)
        return x0, y0, v0

    def step(self, mode, actions, timestep, want_counts=False):
        lib().orc_set_threads(self.threads)
        a = np.ascontiguousarray(actions, dtype=np.int32).reshape(self.E, self.N)
        obs = np.zeros((self.E, self.N, self.R)); rews = np.zeros((self.E, self.N))
        counts = np.zeros((self.E, 2), np.int64) if want_counts else None
        m = MODES[mode] if isinstance(mode, str) else int(mode)
        lib().orc_step(C.byref(self.cfg), C.byref(self._batch), C.c_int(m), _p(a), C.c_int64(timestep),
                       _p(obs), _p(rews), _p(counts))
        return (obs, rews, counts) if want_counts else (obs, rews)

    def obtain_state(self, obs, acts, rews, episode=0, epsilon=1):
        lib().orc_set_threads(self.threads)
        obs = np.ascontiguousarray(obs, dtype=np.float64); rews = np.ascontiguousarray(rews, dtype=np.float64)
        a = np.ascontiguousarray(acts, dtype=np.int32).reshape(self.E, self.N)
        out = np.zeros((self.E, self.N, self.S))
        lib().orc_obtain_state(C.byref(self.cfg), C.byref(self._batch), _p(obs), _p(a), _p(rews),
                               C.c_double(episode), C.c_double(epsilon), _p(out))
        return out

    def information_age(self, timestep):
        out = np.zeros((self.E, 100), np.int32)
        lib().orc_information_age(C.byref(self.cfg), C.byref(self._batch), C.c_int64(timestep), _p(out))
        return out

    def update_velocity(self, draws):
        d = np.ascontiguousarray(draws, dtype=np.int8).reshape(self.E, self.N)
        lib().orc_update_velocity(C.byref(self.cfg), C.byref(self._batch), _p(d))

    # ---- counter-based RNG (specification shared with the CUDA path) ----
    def philox_actions(self, seed, t, env0=0):
        out = np.zeros((self.E, self.N), np.int32)
        lib().orc_philox_actions(C.c_uint64(seed), C.c_int64(env0), C.c_int64(self.E), C.c_int32(self.N),
                                 C.c_int32(self.R), C.c_int64(t), _p(out))
        return out

    def philox_draws(self, seed, episode, env0=0):
        out = np.zeros((self.E, self.N), np.int8)
        lib().orc_philox_draws(C.c_uint64(seed), C.c_int64(env0), C.c_int64(self.E), C.c_int32(self.N),
                               C.c_int64(episode), _p(out))
        return out


def philox(c, k):
    out = (C.c_uint32 * 4)()
    lib().orc_philox(*[C.c_uint32(x) for x in c], *[C.c_uint32(x) for x in k], out)
    return list(out)
