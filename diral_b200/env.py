"""Batched drop-in for the reference's ``TestEnv`` (envs/test_env.py:6-595) on one B200.

``TestEnv(num_envs=E, **EnvironmentTest)`` keeps the reference's constructor kwargs, method names,
argument order and return order, but every array gains a leading env axis and lives on the GPU:

=====================================  =========================================================
reference (one env, Python objects)    here (E envs, torch CUDA tensors)
=====================================  =========================================================
``actions``  list/array [N]            ``[E, N]`` int32
``obs`` dict u -> float[R]             ``[E, N, R]`` float32 (``obs[e][u]`` indexes the same way)
``rews`` float[N]                      ``[E, N]`` float32
``obtain_state`` list of N arrays [S]  ``[E, N, S]`` float32
=====================================  =========================================================

All arithmetic happens in libdiral_env.so (hand-written sm_100a kernels behind the C ABI of
include/diral_env.h).  PyTorch is used for device memory, the current stream and nothing else;
there is no CPU or eager-PyTorch path -- without a CUDA device or the built library the
constructor raises.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib
from ._lib import DiralBuffers, DiralCfg, DiralShaping, MODES, check

IA_BINS = 100          # Network.get_information_age (network.py:566)
METRIC_LEN = 110       # layout documented at diral_episode_metrics in include/diral_env.h
METRIC_FIELDS = ("sum_reward", "sum_collisions", "packets_received", "pairs_in_range", "agent_steps",
                 "bad_actions", "env_slots")

# Network.initialize_mobility_topology_fixed (network.py:81-90), used by reset_mobility_env
_FIXED_TOY = dict(x=[3.0, 5.0, 3.0, 5.0], y=[1.0, 1.0, 2.0, 2.0], v=[0.5, 1.0, 1.25, 1.5])


def cfg_from_kwargs(num_envs: int, env_offset: int, kwargs: dict) -> DiralCfg:
    """``diral_cfg`` from EnvironmentTest-style kwargs with the defaults of TestEnv.__init__
    (test_env.py:12-48).  Unsupported reference variants raise ValueError."""
    st = kwargs.get("State", False)
    if not isinstance(st, dict):
        raise ValueError("State block is required (the reference indexes it unconditionally, test_env.py:27)")
    for key in ("type", "add_reward", "add_action", "add_index", "add_velocity", "action_index", "piggybacking",
                "add_position", "add_positional_dist", "add_positional_dist_piggy", "add_positional_dist_type",
                "num_bins", "add_channel_obs"):
        if key not in st:
            raise ValueError("State block lacks '%s' (test_env.py:27-41 reads it)" % key)
    if st["piggybacking"]:
        raise ValueError("State.piggybacking (observation piggy-backing, test_env.py:71-79) is not enabled by any "
                         "shipped config and is out of scope")
    if kwargs.get("proportional_fair", False):
        raise ValueError("proportional_fair (test_env.py:87-92) is out of scope")
    if st["add_action"] and st["action_index"] not in ("binary", "real"):
        raise ValueError("State.action_index must be 'binary' or 'real' (test_env.py:50-55)")
    c = DiralCfg()
    c.E = int(num_envs); c.env0 = int(env_offset)
    c.N = int(kwargs.get("num_users", 3)); c.R = int(kwargs.get("num_channels", 3)); c.B = int(st["num_bins"])
    c.L = float(kwargs.get("highway_length", 200)); c.C = float(kwargs.get("communication_range", 1))
    c.W = float(kwargs.get("bin_range", 500))
    c.reward_design = int(kwargs.get("reward_design", 1)); c.state_type = int(st["type"])
    c.toy = int(bool(kwargs.get("congestion_test", False)))
    c.mobility = int(bool(kwargs.get("mobility", False)))
    c.mobility_vary = int(bool(kwargs.get("mobility_vary", False)))
    c.design_topology = int(bool(kwargs.get("enable_design_topology", False)))
    c.add_action = int(bool(st["add_action"])); c.action_binary = int(st["action_index"] == "binary")
    c.add_channel_obs = int(bool(st["add_channel_obs"])); c.add_reward = int(bool(st["add_reward"]))
    c.add_index = int(bool(st["add_index"])); c.add_velocity = int(bool(st["add_velocity"]))
    c.add_position = int(bool(st["add_position"])); c.add_positional_dist = int(bool(st["add_positional_dist"]))
    c.add_piggy = int(bool(st["add_positional_dist_piggy"])); c.pos_dist_type = int(st["add_positional_dist_type"])
    c.fingerprint = int(bool(kwargs.get("enable_fingerprint", False)))
    c.age_threshold = 20; c.sentinel = 100000.0
    return c


class Network:
    """The slice of the reference's ``Network`` that callers reach through ``env.network``
    (main_test.py:150 uses ``env.network.get_information_age``)."""

    def __init__(self, env: "TestEnv"):
        self._env = env

    def get_information_age(self, timestep):
        """network.py:560-574, batched: int32 ``[E, 100]``."""
        return self._env.information_age(timestep)

    def get_x_positions(self):
        return self._env.pos_x

    def update_velocity(self, draws=None):
        self._env.update_velocity(draws, force=True)


class TestEnv:
    """Vectorised V2V resource-allocation test simulator (see module docstring)."""

    __test__ = False     # not a pytest class, whatever the name says

    def __init__(self, num_envs=1, device="cuda", seed=0, env_offset=0, init=None, variant="auto",
                 host_format="compact_stream", host_threads=None, **kwargs):
        self.lib = _lib.load()
        self.device = torch.device(device)
        if self.device.type != "cuda" or not torch.cuda.is_available():
            raise RuntimeError("diral_b200.TestEnv runs on a CUDA device only (got device=%r, cuda available=%s); "
                               "there is no CPU path" % (device, torch.cuda.is_available()))
        if self.device.index is None:
            self.device = torch.device("cuda", torch.cuda.current_device())
        self.kwargs = dict(kwargs)
        self.cfg = cfg_from_kwargs(num_envs, env_offset, kwargs)
        self.num_envs = self.E = int(num_envs)
        self.env_offset = int(env_offset)
        self.seed = int(seed)
        # the reference's public attribute names (test_env.py:12-48)
        self.NUM_USERS = self.N = self.cfg.N
        self.NUM_CHANNELS = self.R = self.cfg.R
        self.num_bins = self.B = self.cfg.B
        self.action_space = self.cfg.R
        self.state_space = self.S = int(self.lib.diral_state_space(C.byref(self.cfg)))
        self.mobility = bool(self.cfg.mobility); self.mobility_vary = bool(self.cfg.mobility_vary)
        self.enable_channel = bool(kwargs.get("enable_channel", False))
        self.load_positions = bool(kwargs.get("load_positions", False))
        self.load_file_positions = kwargs.get("load_file_pos", " ")
        self.network = Network(self)

        self._handle = C.c_void_p()
        with torch.cuda.device(self.device):
            check(self.lib.diral_create(C.byref(self.cfg), C.byref(self._handle)))
        if variant != "auto":
            check(self.lib.diral_set_option(self._handle, b"variant", {"group": 1, "block": 2, "block_v1": 3, "row": 4, "pair": 5}[variant]))
        # which kernel runs and how it wants its tables laid out (include/diral_env.h, "Table layouts")
        self.kernel = {1: "group", 2: "block_v1", 3: "row", 4: "pair"}[int(self.lib.diral_get_option(self._handle, b"kernel"))]
        self.layout = int(self.lib.diral_get_option(self._handle, b"layout"))
        self.T = int(self.lib.diral_get_option(self._handle, b"row_stride"))
        self.H = int(self.lib.diral_get_option(self._handle, b"ring_depth"))
        self._dev_calls = True
        self.host_stream = None          # torch.cuda.Stream for step_host / step_host_begin (None: the current stream);
                                         # envs pipelined with step_host_begin / step_host_wait want one each
        self.set_host_format(host_format, host_threads)
        self._alloc()
        self._trace = None
        self._bind()
        self.t = 0             # slot counter of the reset()/step() aliases
        self.episode = 0
        self.reset(seed=seed, init=init)

    # ------------------------------------------------------------------ memory
    def _alloc(self):
        E, N, R, S, dev = self.E, self.N, self.R, self.S, self.device
        f64, i32, f32 = torch.float64, torch.int32, torch.float32
        self.pos_x = torch.zeros((E, N), dtype=f64, device=dev)
        self.pos_y = torch.zeros((E, N), dtype=f64, device=dev)
        self.vel = torch.zeros((E, N), dtype=f64, device=dev)
        self._ring = None
        if self.cfg.add_piggy and self.layout == 1:
            # row layout: _tab_*[e, i, j] is vehicle i's entry about vehicle j (rows padded to T columns); positions
            # live in _ring[e, tick % H, j], entries older than H ticks in the two spill halves of _tab_x
            T = self.T
            self._tab_seq = torch.zeros((E, N, T), dtype=i32, device=dev)
            self._tab_lu = torch.zeros((E, N, T), dtype=i32, device=dev)
            self._tab_x = torch.zeros((2, E, N, T), dtype=f64, device=dev)
            self._ring = torch.zeros((E, self.H, T), dtype=f64, device=dev)
        elif self.cfg.add_piggy:
            # subject-major storage: _tab_*[e, j, i] is vehicle i's entry about vehicle j
            self._tab_seq = torch.zeros((E, N, N), dtype=i32, device=dev)
            self._tab_lu = torch.zeros((E, N, N), dtype=i32, device=dev)
            self._tab_x = torch.zeros((E, N, N), dtype=f64, device=dev)
        else:
            self._tab_seq = self._tab_lu = self._tab_x = None
        self.lat = torch.full((E, N, N), -1, dtype=i32, device=dev)
        self._obs = torch.zeros((E, N, R), dtype=f32, device=dev)
        self._rews = torch.zeros((E, N), dtype=f32, device=dev)
        self._state = torch.zeros((E, N, S), dtype=f32, device=dev)
        self._acc_reward = torch.zeros((E,), dtype=f64, device=dev)
        self._acc_count = torch.zeros((E, 4), dtype=torch.int64, device=dev)
        nscratch = int(self.lib.diral_get_option(self._handle, b"scratch_bytes"))
        self._scratch = torch.zeros((nscratch // 4,), dtype=i32, device=dev) if nscratch else None
        self._metrics = torch.zeros((METRIC_LEN,), dtype=f64, device=dev)
        self._ia = torch.zeros((E, IA_BINS), dtype=i32, device=dev)
        self._actions = torch.zeros((E, N), dtype=i32, device=dev)

    def _bind(self):
        b = DiralBuffers()
        ptr = lambda t: t.data_ptr() if t is not None else None
        b.pos_x, b.pos_y, b.vel = ptr(self.pos_x), ptr(self.pos_y), ptr(self.vel)
        b.tab_seq, b.tab_lu, b.tab_x = ptr(self._tab_seq), ptr(self._tab_lu), ptr(self._tab_x)
        b.lat, b.obs, b.rews, b.state = ptr(self.lat), ptr(self._obs), ptr(self._rews), ptr(self._state)
        b.acc_reward, b.acc_count, b.scratch = ptr(self._acc_reward), ptr(self._acc_count), ptr(self._scratch)
        b.trace = ptr(self._trace)
        b.trace_len = 0 if self._trace is None else int(self._trace.shape[0])
        b.ring = ptr(self._ring)
        check(self.lib.diral_bind(self._handle, C.byref(b)))

    def close(self):
        if getattr(self, "_handle", None) is not None and self._handle.value:
            self.lib.diral_destroy(self._handle)
            self._handle = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_host_format(self, host_format="compact_stream", host_threads=None, shared_pool=False):
        """How ``step_host`` moves a slot's results to the host: ``"full"`` copies the [E, N, S] float32 rows over
        PCIe; ``"compact"`` copies only what the host cannot know (VPD bin counts as bytes, rewards, ...) and lets
        ``host_threads`` library threads assemble the same rows in the caller's buffer (include/diral_env.h);
        ``"compact_stream"`` is the same record written by ONE launch straight into mapped host memory, the kernel
        raising a flag per chunk of environments that releases the assembly threads (lane-group kernel).
        ``host_threads=None``: the CPUs this process may use, shared between the ranks of a torchrun job.
        ``shared_pool=True``: this env's rows are assembled by the process-wide pool (sized by the first env that asks
        for it) -- what several envs pipelined with ``step_host_begin`` / ``step_host_wait`` should use."""
        formats = {"full": 0, "compact": 1, "compact_zero_copy": 2, "compact_stream": 3}
        if host_format not in formats:
            raise ValueError("host_format must be one of %s" % ", ".join(repr(k) for k in formats))
        check(self.lib.diral_set_option(self._handle, b"host_format", formats[host_format]))
        if host_threads is None:
            import os
            cpus = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
            local = max(int(os.environ.get("LOCAL_WORLD_SIZE", "1")), 1)
            share = cpus // local
            host_threads = max(min(share - (2 if share >= 8 else 1), 32), 1)   # room for the caller and the CUDA runtime's threads
        check(self.lib.diral_set_option(self._handle, b"host_threads", int(host_threads)))
        check(self.lib.diral_set_option(self._handle, b"host_pool_shared", int(bool(shared_pool))))
        self.host_format = host_format if self.lib.diral_get_option(self._handle, b"compact_ok") else "full"

    # ------------------------------------------------------------------ helpers
    def _stream(self):
        self._dev_calls = True
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def _host_stream(self):
        # the host-buffer calls return only after their own work has been synchronised, so they may run on a private
        # stream; everything that hands device tensors back stays on torch's current stream (and is waited for here)
        if self.host_stream is not None:
            if self._dev_calls:
                self.host_stream.wait_stream(torch.cuda.current_stream(self.device))
                self._dev_calls = False
            return C.c_void_p(self.host_stream.cuda_stream)
        return self._stream()

    def _as_actions(self, actions):
        if isinstance(actions, torch.Tensor):
            a = actions
            if a.device != self.device or a.dtype != torch.int32 or not a.is_contiguous():
                a = a.to(device=self.device, dtype=torch.int32).contiguous()
        else:
            a = torch.as_tensor(np.ascontiguousarray(actions, dtype=np.int32), device=self.device)
        if a.numel() != self.E * self.N:
            raise ValueError("actions must hold num_envs*num_users = %d entries (got shape %s)"
                             % (self.E * self.N, tuple(a.shape)))
        return a.view(self.E, self.N)

    def _as_f64(self, v, what):
        if v is None:
            return None
        t = torch.as_tensor(v, dtype=torch.float64).to(self.device)
        if t.numel() == self.N:
            t = t.reshape(1, self.N).expand(self.E, self.N)
        if t.numel() != self.E * self.N and tuple(t.shape) != (self.E, self.N):
            raise ValueError("%s must have shape [%d, %d] or [%d]" % (what, self.E, self.N, self.N))
        return t.reshape(self.E, self.N).contiguous()

    # ------------------------------------------------------------------ reference surface
    def reset(self, seed=None, init=None):
        """TestEnv.__init__ / Network.__init__ (test_env.py:98-101, network.py:15-119): zero tables,
        ``last_arrival_time = -1`` and a new topology -- ``init=(x0, y0, v0)`` arrays of shape
        [E, N] or [N], else the counter-based generator keyed by (seed, global env index).
        Returns the state for all-zero observations, like the first ``obtain_state`` would."""
        if seed is not None:
            self.seed = int(seed)
        with torch.cuda.device(self.device):
            if init is not None:
                x0, y0, v0 = (self._as_f64(v, n) for v, n in zip(init, ("x0", "y0", "v0")))
                check(self.lib.diral_reset(self._handle, x0.data_ptr(), y0.data_ptr(), v0.data_ptr(),
                                           C.c_uint64(self.seed), self._stream()))
            else:
                check(self.lib.diral_reset(self._handle, None, None, None, C.c_uint64(self.seed), self._stream()))
        self.t = 0
        self.episode = 0
        self._obs.zero_(); self._rews.zero_()
        if hasattr(self, "_shape_sums"):          # shaping state of main_test.py:48-56,73 starts over too
            self._sum_ia_prev.zero_(); self._ia_counter.zero_(); self._prev_actions.fill_(-1)
        return self

    def sample(self, t=None):
        """test_env.py:116-122: uniform random actions, ``[E, N]`` int32 (Philox, reproducible per
        (seed, global env index, vehicle, t))."""
        out = torch.empty((self.E, self.N), dtype=torch.int32, device=self.device)
        with torch.cuda.device(self.device):
            check(self.lib.diral_sample(self._handle, C.c_uint64(self.seed), C.c_int64(self.t if t is None else int(t)),
                                        out.data_ptr(), self._stream()))
        return out

    def _step(self, mode, actions, timestep, build_state=False, episode_number=0, epsilon=1):
        a = self._as_actions(actions) if actions is not None else None
        with torch.cuda.device(self.device):
            check(self.lib.diral_step(self._handle, MODES[mode], a.data_ptr() if a is not None else None,
                                      C.c_int64(int(timestep)), int(bool(build_state)), float(episode_number),
                                      float(epsilon), C.c_uint64(self.seed),
                                      self._actions.data_ptr() if a is None else None, self._stream()))
        return self._obs, self._rews

    def my_step(self, actions, timestep):
        """test_env.py:124-266 -> ``(obs [E,N,R], rews [E,N])`` (views, valid until the next step)."""
        return self._step("my_step", actions, timestep)

    def my_step_design(self, actions, timestep):
        """test_env.py:269-316."""
        return self._step("my_step_design", actions, timestep)

    def my_step_ch(self, actions, time_step):
        """test_env.py:351-443 (PRR reward)."""
        return self._step("my_step_ch", actions, time_step)

    def obtain_state(self, obs, acts, rewards, episode_number=0, epsilon=1):
        """test_env.py:527-583 -> ``[E, N, S]`` float32 (a view, valid until the next call)."""
        a = self._as_actions(acts)
        o = torch.as_tensor(obs, dtype=torch.float32, device=self.device).reshape(self.E, self.N, self.R).contiguous()
        r = torch.as_tensor(rewards, dtype=torch.float32, device=self.device).reshape(self.E, self.N).contiguous()
        with torch.cuda.device(self.device):
            check(self.lib.diral_obtain_state(self._handle, o.data_ptr(), a.data_ptr(), r.data_ptr(),
                                              float(episode_number), float(epsilon), self._state.data_ptr(),
                                              self._stream()))
        return self._state

    def get_x_pos(self):
        return self.pos_x

    def reset_mobility_env(self):
        """test_env.py:479-484 -> Network.reset_positions (network.py:181-187): the fixed 4-vehicle toy
        topology with fresh tables.  Only the vehicles are rebuilt, as in the reference: ``last_arrival_time``,
        the slot / episode counters, the episode accumulators and the last obs / rewards keep their values."""
        if self.N != 4:
            raise ValueError("reset_mobility_env installs the fixed 4-vehicle topology (network.py:81-90)")
        x0, y0, v0 = (self._as_f64(_FIXED_TOY[k], k) for k in ("x", "y", "v"))
        with torch.cuda.device(self.device):
            check(self.lib.diral_reset_topology(self._handle, x0.data_ptr(), y0.data_ptr(), v0.data_ptr(),
                                                C.c_uint64(self.seed), self._stream()))

    def get_total_users(self):
        return self.NUM_USERS

    def get_num_ch(self):
        return self.NUM_CHANNELS

    def get_state_space(self):
        return self.state_space

    def get_action_space(self):
        return self.action_space

    def update_velocity(self, draws=None, force=False, episode=None):
        """test_env.py:498-504 -> network.py:208-222; ``draws`` [E, N] in {1,2,3} replays recorded
        ``random.randrange(1, 4)`` outcomes, else Philox keyed by (seed, global env, vehicle, episode).  The
        reference calls this once per episode (main_test.py:226-236), so every call advances ``self.episode``
        -- two calls never repeat a draw; ``episode=`` pins the key explicitly."""
        if not (self.mobility_vary or force):
            return
        ep = self.episode if episode is None else int(episode)
        d = None
        if draws is not None:
            d = torch.as_tensor(np.ascontiguousarray(draws, dtype=np.int8), device=self.device).reshape(self.E, self.N)
        with torch.cuda.device(self.device):
            check(self.lib.diral_update_velocity(self._handle, d.data_ptr() if d is not None else None,
                                                 C.c_uint64(self.seed), C.c_int64(ep), self._stream()))
        self.episode = ep + 1

    def load_saved_positions(self, trace=None):
        """test_env.py:109-114 -> Network.load_x_positions (network.py:171-178): replay a [T, N]
        trace of x positions (``load_file_pos`` .npy, or an array passed directly)."""
        if trace is None:
            if not self.load_positions:
                return
            trace = np.load(self.load_file_positions)
        tr = np.ascontiguousarray(trace, dtype=np.float64)
        if tr.ndim != 2 or tr.shape[1] != self.N:
            raise ValueError("trace must have shape [T, %d]" % self.N)
        self._trace = torch.as_tensor(tr, device=self.device)
        self._bind()

    # ------------------------------------------------------------------ gym-style aliases (fused path)
    def step(self, actions=None, episode_number=0, epsilon=1):
        """One slot through the FUSED kernel: ``my_step*`` (mode from ``enable_channel``,
        main_test.py:143-146) + ``obtain_state`` in one pass over the tables.
        ``actions=None`` draws them on device.  Returns ``(state, rews, info)``."""
        mode = "my_step_ch" if self.enable_channel else "my_step"
        self._step(mode, actions, self.t, True, episode_number, epsilon)
        self.t += 1
        return self._state, self._rews, {"obs": self._obs, "actions": self._actions if actions is None else actions}

    def rollout(self, T, mode=None):
        """T fused slots with on-device actions (throughput runs)."""
        mode = mode or ("my_step_ch" if self.enable_channel else "my_step")
        with torch.cuda.device(self.device):
            check(self.lib.diral_rollout(self._handle, MODES[mode], int(T), C.c_int64(self.t), C.c_uint64(self.seed),
                                         self._stream()))
        self.t += int(T)
        return self._state, self._rews

    def step_host(self, h_actions, h_state, h_rews, h_obs=None, mode=None, episode_number=0, epsilon=1):
        """Host-buffer step (diral_step_host): numpy/pinned-tensor actions in, state/rews out."""
        mode = mode or ("my_step_ch" if self.enable_channel else "my_step")
        ptr = lambda t: t.data_ptr() if isinstance(t, torch.Tensor) else t.ctypes.data
        with torch.cuda.device(self.device):
            check(self.lib.diral_step_host(self._handle, MODES[mode], ptr(h_actions), C.c_int64(self.t),
                                           float(episode_number), float(epsilon), ptr(h_state), ptr(h_rews),
                                           ptr(h_obs) if h_obs is not None else None, self._host_stream()))
        self.t += 1

    def step_host_begin(self, h_actions, h_state, h_rews, mode=None, episode_number=0, epsilon=1):
        """First half of ``step_host`` (diral_step_host_begin, ``host_format="compact_stream"`` on a lane-group
        configuration): enqueues the slot and returns; ``step_host_wait`` returns once ``h_state`` / ``h_rews`` hold
        it.  Two or more envs stepped this way overlap one env's kernel and PCIe records with another's row assembly
        (give each its own ``host_threads`` share)."""
        mode = mode or ("my_step_ch" if self.enable_channel else "my_step")
        ptr = lambda t: t.data_ptr() if isinstance(t, torch.Tensor) else t.ctypes.data
        check(self.lib.diral_step_host_begin(self._handle, MODES[mode], ptr(h_actions), C.c_int64(self.t),
                                             float(episode_number), float(epsilon), ptr(h_state), ptr(h_rews),
                                             self._host_stream()))
        self.t += 1

    def step_host_wait(self):
        check(self.lib.diral_step_host_wait(self._handle))

    # ------------------------------------------------------------------ metrics
    def information_age(self, timestep):
        with torch.cuda.device(self.device):
            check(self.lib.diral_information_age(self._handle, C.c_int64(int(timestep)), self._ia.data_ptr(),
                                                 self._stream()))
        return self._ia

    def episode_metrics(self, timestep=None):
        """Per-device end-of-episode metric vector (float64[110], layout in include/diral_env.h);
        clears the accumulators.  diral_b200.dist.all_reduce_metrics sums it across GPUs."""
        with torch.cuda.device(self.device):
            check(self.lib.diral_episode_metrics(self._handle, C.c_int64(self.t if timestep is None else int(timestep)),
                                                 self._metrics.data_ptr(), self._stream()))
        return self._metrics

    def shape_rewards(self, actions, rewards, timestep, ia_averaging=False, ia_penalty_enable=False,
                      ia_penalty_threshold=5, ia_penalty_value=-10, global_reward_avg=False):
        """The per-slot caller epilogue of the reference's driver (main_test.py:150-206, utils/misc.py):
        information age of this slot, its weighted sum, and the reward shaping options, applied IN PLACE
        to ``rewards`` ([E, N] float32, e.g. the tensor ``my_step*`` returned).  Call it after
        ``obtain_state`` -- the reference builds the state from the unshaped rewards (main_test.py:164).
        Returns ``(rewards, slot_sums [E, 3] = (sum_r, collisions, weighted IA), ia [E, 100])``."""
        if not hasattr(self, "_shape_sums"):
            dev = self.device
            self._shape_sums = torch.zeros((self.E, 3), dtype=torch.float64, device=dev)
            self._sum_ia_prev = torch.zeros((self.E,), dtype=torch.int64, device=dev)        # main_test.py:73
            self._ia_counter = torch.zeros((self.E, self.N), dtype=torch.int32, device=dev)  # main_test.py:55
            self._prev_actions = torch.full((self.E, self.N), -1, dtype=torch.int32, device=dev)   # :56
        a = self._as_actions(actions)
        if not (isinstance(rewards, torch.Tensor) and rewards.dtype == torch.float32 and rewards.is_contiguous()
                and rewards.device == self.device and rewards.numel() == self.E * self.N):
            raise ValueError("rewards must be a contiguous float32 tensor of num_envs*num_users entries on %s" % self.device)
        cfg = DiralShaping(int(bool(ia_averaging)), int(bool(ia_penalty_enable)), int(ia_penalty_threshold),
                           int(bool(global_reward_avg)), float(ia_penalty_value))
        with torch.cuda.device(self.device):
            check(self.lib.diral_shape_rewards(self._handle, C.byref(cfg), a.data_ptr(), C.c_int64(int(timestep)),
                                               rewards.data_ptr(), self._sum_ia_prev.data_ptr(), self._ia_counter.data_ptr(),
                                               self._prev_actions.data_ptr(), self._shape_sums.data_ptr(),
                                               self._ia.data_ptr(), self._stream()))
        return rewards, self._shape_sums, self._ia

    def launch_count(self):
        return int(self.lib.diral_launch_count(self._handle))

    # ------------------------------------------------------------------ reference-layout views
    def _need_tables(self):
        if self._tab_seq is None:
            raise AttributeError("neighbour tables exist only with State.add_positional_dist_piggy (vehicle.py:20-33 "
                                 "are never read otherwise)")

    @property
    def tab_seq(self):
        """``[E, i, j]`` = vehicles[i].pos_of_neighbors[j]["seq_number"] (vehicle.py:32)."""
        self._need_tables()
        return self._tab_seq[:, :, :self.N] if self.layout == 1 else self._tab_seq.transpose(1, 2)

    @property
    def tab_lu(self):
        self._need_tables()
        return self._tab_lu[:, :, :self.N] if self.layout == 1 else self._tab_lu.transpose(1, 2)

    @property
    def tab_x(self):
        """``[E, i, j]`` = vehicles[i].pos_of_neighbors[j]["xpos"] (vehicle.py:30).  The row layout keeps positions in a
        ring by tick, so the table is rebuilt on demand (diral_materialize_x)."""
        self._need_tables()
        if self.layout == 0:
            return self._tab_x.transpose(1, 2)
        out = torch.empty((self.E, self.N, self.N), dtype=torch.float64, device=self.device)
        with torch.cuda.device(self.device):
            check(self.lib.diral_materialize_x(self._handle, out.data_ptr(), self._stream()))
        return out

    @property
    def tab_y(self):
        """ypos is derived: pos_y never changes between resets, so an entry that has ever been written
        (seq > 0) holds pos_y of its subject (vehicle.py:31,43,60)."""
        self._need_tables()
        return torch.where(self.tab_seq > 0, self.pos_y[:, None, :].expand(-1, self.N, -1),
                           torch.zeros((), dtype=torch.float64, device=self.device))

    # ------------------------------------------------------------------ checkpoint
    _STATE_TENSORS = ("pos_x", "pos_y", "vel", "lat", "_tab_seq", "_tab_lu", "_tab_x", "_ring", "_acc_reward", "_acc_count",
                      "_obs", "_rews", "_state")

    def state_dict(self):
        """Everything a restored env needs to continue bit for bit: kinematics, tables, last_arrival_time, the
        episode accumulators, the last outputs, and the host-side counters the kernels derive keys from."""
        d = {k.lstrip("_"): getattr(self, k).clone() for k in self._STATE_TENSORS if getattr(self, k) is not None}
        d.update(t=self.t, episode=self.episode, seed=self.seed,
                 ticks=int(self.lib.diral_get_option(self._handle, b"ticks")),
                 lat_live=int(self.lib.diral_get_option(self._handle, b"lat_live")),
                 track_lat=int(self.lib.diral_get_option(self._handle, b"track_lat")))
        if hasattr(self, "_shape_sums"):
            d.update(sum_ia_prev=self._sum_ia_prev.clone(), ia_counter=self._ia_counter.clone(),
                     prev_actions=self._prev_actions.clone())
        return d

    def load_state_dict(self, d):
        """Inverse of ``state_dict`` on an env of the same configuration."""
        for k in self._STATE_TENSORS:
            dst = getattr(self, k)
            if dst is None:
                continue
            src = d[k.lstrip("_")]
            if tuple(src.shape) != tuple(dst.shape):
                raise ValueError("state_dict['%s'] has shape %s, this env needs %s" % (k.lstrip("_"), tuple(src.shape), tuple(dst.shape)))
            dst.copy_(src)
        self.t, self.episode, self.seed = int(d["t"]), int(d["episode"]), int(d.get("seed", self.seed))
        for name in ("ticks", "lat_live", "track_lat"):
            check(self.lib.diral_set_option(self._handle, name.encode(), int(d[name])))
        if "sum_ia_prev" in d:
            if not hasattr(self, "_shape_sums"):     # the shaping state is created lazily (shape_rewards)
                dev = self.device
                self._shape_sums = torch.zeros((self.E, 3), dtype=torch.float64, device=dev)
                self._sum_ia_prev = torch.zeros((self.E,), dtype=torch.int64, device=dev)
                self._ia_counter = torch.zeros((self.E, self.N), dtype=torch.int32, device=dev)
                self._prev_actions = torch.full((self.E, self.N), -1, dtype=torch.int32, device=dev)
            self._sum_ia_prev.copy_(d["sum_ia_prev"]); self._ia_counter.copy_(d["ia_counter"])
            self._prev_actions.copy_(d["prev_actions"])
        return self


BatchedTestEnv = TestEnv
