#!/usr/bin/env python
"""Generate golden input/output vectors by executing the UNMODIFIED reference env.

Run in the build container only (needs /root/reference, which never travels to the
GPU box):

    python tests/golden/make_golden.py

It imports ``envs/test_env.py`` from the reference with the two shims SURVEY.md
section 8(c) lists (``sys.path`` += envs/ for the py2 implicit imports, and a stub
``matplotlib`` because ``network.py:6`` imports it for the dead ``plot_fc``), drives
``TestEnv`` through ``my_step`` / ``my_step_design`` / ``my_step_ch`` + ``obtain_state``
(reference ``envs/test_env.py:124,269,351,527``) with recorded actions, and stores
every input and output as ``tests/golden/<case>.npz``.  The fixtures are what pins
``oracle/`` (and, through it, the CUDA path) to the reference.

Nothing in here is product code and nothing is copied from the reference: the
reference is only *executed*.
"""
from __future__ import annotations

import contextlib
import io
import os
import random
import sys
import types

import numpy as np

REF = os.environ.get("DIRAL_REFERENCE", "/root/reference")
OUT = os.path.dirname(os.path.abspath(__file__))


def load_reference():
    """Import the reference TestEnv with the two shims (no source edits)."""
    if "matplotlib" not in sys.modules:
        mpl = types.ModuleType("matplotlib")
        plt = types.ModuleType("matplotlib.pyplot")
        mpl.pyplot = plt
        sys.modules["matplotlib"] = mpl
        sys.modules["matplotlib.pyplot"] = plt
    envs = os.path.join(REF, "envs")
    if envs not in sys.path:
        sys.path.insert(0, envs)
    import test_env  # noqa: E402  (the reference module)
    return test_env.TestEnv


def shipped_state(**over):
    """``State`` block exactly as the shipped YAML has it
    (configs/4ue_3r_toy/*_dis_03.yaml:57-71)."""
    st = dict(type=2, add_action=True, add_reward=False, add_index=False,
              add_velocity=False, action_index="binary", piggybacking=False,
              add_position=False, add_positional_dist=False,
              add_positional_dist_piggy=True, add_positional_dist_type=2,
              add_channel_obs=False, num_bins=20)
    st.update(over)
    return st


def dump_tables(env):
    n = env.NUM_USERS
    xpos = np.zeros((n, n), np.float64)
    ypos = np.zeros((n, n), np.float64)
    seq = np.zeros((n, n), np.int64)
    lu = np.zeros((n, n), np.int64)
    for i in range(n):
        tab = env.network.vehicles[i].pos_of_neighbors
        for j in range(n):
            xpos[i, j] = tab[j]["xpos"]
            ypos[i, j] = tab[j]["ypos"]
            seq[i, j] = tab[j]["seq_number"]
            lu[i, j] = tab[j]["last_updated"]
    return xpos, ypos, seq, lu


def dump_lat(env):
    n = env.NUM_USERS
    lat = np.zeros((n, n), np.int64)
    for t in range(n):
        for r in range(n):
            lat[t, r] = env.network.last_arrival_time[t][r]
    return lat


def run_case(name, kwargs, modes, T, seed, *, fixed_toy=False, episode_interval=0,
             trace=None, actions=None, fingerprint=False, use_sample=False):
    """Run one reference env for T slots; ``modes`` is a per-slot list (or one str)."""
    TestEnv = load_reference()
    if isinstance(modes, str):
        modes = [modes] * T
    np.random.seed(seed)
    random.seed(seed)
    kwargs = dict(kwargs)
    if trace is not None:
        kwargs["load_positions"] = True
        kwargs["load_file_pos"] = trace
    with contextlib.redirect_stdout(io.StringIO()):
        env = TestEnv(**kwargs)
        if fixed_toy:
            env.reset_mobility_env()
        if trace is not None:
            env.load_saved_positions()
    n, r = env.NUM_USERS, env.NUM_CHANNELS
    veh = env.network.vehicles
    x0 = np.array([float(v.pos_x) for v in veh])
    y0 = np.array([float(v.pos_y) for v in veh])
    v0 = np.array([float(v.velocity) for v in veh])

    rs = np.random.RandomState(seed + 77)
    acts_l, obs_l, rew_l, st_l, x_l, v_l, ia_l = [], [], [], [], [], [], []
    tx_l, ty_l, ts_l, tl_l, lat_l, draws_l = [], [], [], [], [], []

    draws_now = []
    real_randrange = random.randrange

    def logging_randrange(*a, **k):
        d = real_randrange(*a, **k)
        draws_now.append(d)
        return d

    for t in range(T):
        if actions is not None:
            a = np.asarray(actions[t], dtype=np.int64)
        elif use_sample:
            a = np.asarray(env.sample(), dtype=np.int64)
        else:
            a = rs.randint(0, r, size=n).astype(np.int64)
        mode = modes[t]
        with contextlib.redirect_stdout(io.StringIO()):
            obs, rews = getattr(env, mode)(a, t)
        rews = np.array(rews, dtype=np.float64)
        ep, eps = (t // 25, 0.9992 ** (t // 25)) if fingerprint else (0, 1)
        with contextlib.redirect_stdout(io.StringIO()):
            state = env.obtain_state(obs, a, rews, ep, eps)
        acts_l.append(a)
        obs_l.append(np.stack([np.asarray(obs[u], dtype=np.float64) for u in range(n)]))
        rew_l.append(rews)
        st_l.append(np.stack([np.asarray(s, dtype=np.float64) for s in state]))
        x_l.append(np.array([float(v.pos_x) for v in veh], dtype=np.float64))
        ia_l.append(np.array(env.network.get_information_age(t), dtype=np.int64))
        tx, ty, ts, tl = dump_tables(env)
        tx_l.append(tx); ty_l.append(ty); ts_l.append(ts); tl_l.append(tl)
        lat_l.append(dump_lat(env))
        draws_now = []
        if episode_interval and t % episode_interval == episode_interval - 1:
            random.randrange = logging_randrange
            try:
                env.update_velocity()
            finally:
                random.randrange = real_randrange
        d = np.array(draws_now if draws_now else [0] * n, dtype=np.int8)
        if len(d) != n:
            d = np.zeros(n, np.int8)
        draws_l.append(d)
        v_l.append(np.array([float(v.velocity) for v in veh], dtype=np.float64))

    st = kwargs["State"]
    meta = dict(
        num_users=n, num_channels=r, num_bins=int(st["num_bins"]),
        highway_length=float(kwargs.get("highway_length", 200)),
        communication_range=float(kwargs.get("communication_range", 1)),
        bin_range=float(kwargs.get("bin_range", 500)),
        reward_design=int(kwargs.get("reward_design", 1)),
        congestion_test=bool(kwargs.get("congestion_test", False)),
        mobility=bool(kwargs.get("mobility", False)),
        mobility_vary=bool(kwargs.get("mobility_vary", False)),
        enable_design_topology=bool(kwargs.get("enable_design_topology", False)),
        enable_fingerprint=bool(kwargs.get("enable_fingerprint", False)),
        state_type=int(st["type"]), add_action=bool(st["add_action"]),
        action_binary=(st["action_index"] == "binary"),
        add_channel_obs=bool(st["add_channel_obs"]), add_reward=bool(st["add_reward"]),
        add_index=bool(st["add_index"]), add_velocity=bool(st["add_velocity"]),
        add_position=bool(st["add_position"]),
        add_positional_dist=bool(st["add_positional_dist"]),
        add_positional_dist_piggy=bool(st["add_positional_dist_piggy"]),
        add_positional_dist_type=int(st["add_positional_dist_type"]),
        state_space=int(env.get_state_space()), episode_interval=int(episode_interval),
        fingerprint_args=bool(fingerprint),
    )
    arrays = dict(
        x0=x0, y0=y0, v0=v0,
        modes=np.array(modes), actions=np.stack(acts_l).astype(np.int32),
        obs=np.stack(obs_l), rews=np.stack(rew_l), state=np.stack(st_l),
        pos_x=np.stack(x_l), vel=np.stack(v_l), ia=np.stack(ia_l).astype(np.int32),
        tab_x=np.stack(tx_l), tab_y=np.stack(ty_l),
        tab_seq=np.stack(ts_l).astype(np.int32), tab_lu=np.stack(tl_l).astype(np.int32),
        lat=np.stack(lat_l).astype(np.int32), draws=np.stack(draws_l),
    )
    if trace is not None:
        arrays["trace"] = np.load(trace)
    for k, v in meta.items():
        arrays["meta_" + k] = np.array(v)
    path = os.path.join(OUT, name + ".npz")
    np.savez_compressed(path, **arrays)
    print("%-34s T=%-3d N=%-3d R=%-3d S=%-3d  %6.1f KB" % (
        name, T, n, r, meta["state_space"], os.path.getsize(path) / 1024.0))


def main():
    toy = dict(congestion_test=True, load_positions=False, num_channels=3, num_users=4,
               mobility=True, mobility_vary=False, highway_length=100,
               enable_fingerprint=False, reward_design=2, communication_range=250)
    design6 = dict(congestion_test=False, num_channels=5, num_users=6, mobility=True,
                   highway_length=1200, reward_design=2, communication_range=250,
                   bin_range=500, enable_design_topology=True)

    # ---- SURVEY.md KAT-1..5 -------------------------------------------------------
    kat_actions = [[0, 1, 2, 3, 4, 0], [0, 0, 1, 2, 3, 4], [1, 1, 1, 2, 3, 4]]
    run_case("kat1_design6_step_design", dict(design6, State=shipped_state()),
             "my_step_design", 3, 1, actions=kat_actions)
    run_case("kat2_design6_my_step", dict(design6, State=shipped_state(add_channel_obs=True)),
             "my_step", 3, 1, actions=[[0, 1, 2, 3, 4, 0], [0, 0, 1, 2, 3, 4], [0, 1, 0, 2, 3, 4]])
    for d in (2, 3, 4):
        run_case("kat3_design6_ch_d%d" % d, dict(design6, reward_design=d, State=shipped_state()),
                 "my_step_ch", 4, 1,
                 actions=[[0, 1, 2, 3, 4, 0], [0, 0, 1, 2, 3, 4], [0, 1, 0, 2, 3, 4], [4, 3, 2, 1, 0, 0]])
    run_case("kat4_toy_fixed", dict(toy, State=shipped_state()), "my_step", 5, 1, fixed_toy=True,
             actions=[[0, 1, 2, 0], [2, 1, 0, 2], [1, 1, 1, 1], [0, 0, 1, 1], [2, 0, 0, 1]])
    run_case("kat5_toy_seed0_sample", dict(toy, State=shipped_state()), "my_step", 30, 0,
             use_sample=True)

    # ---- shipped toy config, long enough to leave the 20-slot phantom phase -------
    run_case("toy4x3_shipped_T80", dict(toy, State=shipped_state()), "my_step", 80, 11)
    # main_test.py order: my_step once, my_step_design warm-up, then my_step
    sched = ["my_step"] + ["my_step_design"] * 20 + ["my_step"] * 30
    run_case("toy4x3_pretrain_schedule", dict(toy, State=shipped_state()), sched, len(sched), 12)

    # ---- congested demo 6x5, random topology ----------------------------------------
    c2 = dict(congestion_test=False, num_channels=5, num_users=6, mobility=True,
              highway_length=1170, reward_design=2, communication_range=250, bin_range=500)
    run_case("c2_6x5_my_step", dict(c2, State=shipped_state()), "my_step", 60, 21)
    run_case("c2_6x5_design_topo_T1300", dict(design6, State=shipped_state()), "my_step", 130, 22)

    # ---- headline shape 32x20 ---------------------------------------------------------
    c3 = dict(congestion_test=False, num_channels=20, num_users=32, mobility=True,
              highway_length=800, reward_design=2, communication_range=250, bin_range=500)
    run_case("c3_32x20_my_step", dict(c3, State=shipped_state()), "my_step", 40, 31)
    run_case("c3_32x20_my_step_chanobs", dict(c3, State=shipped_state(add_channel_obs=True)),
             "my_step", 30, 32)
    for d in (2, 3, 4):
        run_case("c3_32x20_ch_d%d" % d, dict(c3, reward_design=d, State=shipped_state()),
                 "my_step_ch", 30, 33 + d)
    run_case("c3_32x20_step_design", dict(c3, State=shipped_state()), "my_step_design", 30, 38)
    # sparse: long highway, most vehicles out of range (sentinel 100000, partial tables)
    run_case("sparse_24x6_L6000", dict(c3, num_users=24, num_channels=6, highway_length=6000,
                                       State=shipped_state(add_channel_obs=True)), "my_step", 60, 39)
    run_case("sparse_24x6_L6000_ch", dict(c3, num_users=24, num_channels=6, highway_length=6000,
                                          reward_design=2, State=shipped_state()), "my_step_ch", 60, 40)

    # ---- larger-than-a-warp agent counts ----------------------------------------------
    run_case("n48x10_my_step", dict(c3, num_users=48, num_channels=10, highway_length=1200,
                                    State=shipped_state()), "my_step", 25, 41)
    run_case("n70x16_ch_d3", dict(c3, num_users=70, num_channels=16, highway_length=2500,
                                  reward_design=3, State=shipped_state()), "my_step_ch", 12, 42)
    # beyond 128 vehicles: packed keys only in shared memory, 1024-thread CTAs, positions gathered from global memory
    run_case("n130x40_my_step", dict(c3, num_users=130, num_channels=40, highway_length=3250,
                                     State=shipped_state()), "my_step", 8, 43)

    # ---- every reward design, collision-heavy (8 UE x 3 res), toy and non-toy -----------
    for d in (1, 2, 3, 4, 5):
        for toyflag in (False, True):
            kw = dict(congestion_test=toyflag, num_channels=3, num_users=8, mobility=True,
                      highway_length=(60 if toyflag else 700), reward_design=d,
                      communication_range=(250 if toyflag else 150), bin_range=500)
            run_case("rd%d_%s_8x3" % (d, "toy" if toyflag else "net"),
                     dict(kw, State=shipped_state()), "my_step", 40, 50 + d)

    # ---- every state-vector block -----------------------------------------------------
    full = shipped_state(add_channel_obs=True, add_reward=True, add_index=True, add_velocity=True,
                         add_position=True, add_positional_dist=True)
    run_case("state_all_blocks_12x4", dict(c3, num_users=12, num_channels=4, highway_length=900,
                                           enable_fingerprint=True, State=full),
             "my_step", 40, 61, fingerprint=True)
    run_case("state_real_action_vpd1_12x4",
             dict(c3, num_users=12, num_channels=4, highway_length=900,
                  State=shipped_state(action_index="real", add_positional_dist_type=1, add_index=True)),
             "my_step", 40, 62)
    run_case("state_vpd1_bins10_16x5",
             dict(c3, num_users=16, num_channels=5, highway_length=700,
                  State=shipped_state(add_positional_dist_type=1, num_bins=10)), "my_step_ch", 40, 63)
    run_case("state_bins40_W200_10x4",
             dict(c3, num_users=10, num_channels=4, highway_length=500, bin_range=200,
                  State=shipped_state(num_bins=40)), "my_step", 40, 64)
    run_case("state_no_piggy_direct_10x4",
             dict(c3, num_users=10, num_channels=4, highway_length=600,
                  State=shipped_state(add_positional_dist_piggy=False, add_positional_dist=True,
                                      add_channel_obs=True)), "my_step", 30, 65)

    # ---- mobility variants ---------------------------------------------------------------
    run_case("mobility_vary_10x4", dict(c3, num_users=10, num_channels=4, highway_length=400,
                                        mobility_vary=True, State=shipped_state(add_velocity=True)),
             "my_step", 110, 71, episode_interval=25)
    rs = np.random.RandomState(5)
    trace = np.cumsum(rs.uniform(0.5, 3.0, size=(17, 6)), axis=0) % 1170.0
    tpath = "/tmp/diral_trace_6.npy"
    np.save(tpath, trace)
    run_case("trace_replay_6x5", dict(c2, State=shipped_state(add_position=True)), "my_step", 40, 72,
             trace=tpath)
    run_case("static_design_topo_6x5", dict(design6, mobility=False, State=shipped_state()),
             "my_step_ch", 30, 73)


if __name__ == "__main__":
    main()
