"""GPU parity, second batch: full-size comparisons against the oracle at the BASELINE shapes, the device topology
presets against the reference's fixtures, the compact host format of diral_step_host against the full one, the
checkpoint round trip, and the kernel-selection edges ("parity gaps").

Same bars as tests/test_gpu_parity.py: integer state and float64 positions bit-exact, float32 outputs equal to the
float32 rounding of the oracle's float64 values.
"""
import numpy as np
import pytest
import torch

from golden_util import kwargs_from_meta, load_golden

pytestmark = pytest.mark.gpu


def _env(E, variant="auto", **kw):
    from diral_b200 import TestEnv
    return TestEnv(num_envs=E, device="cuda", variant=variant, **kw)


def _np(t):
    return t.detach().cpu().numpy()


def _state(**over):
    st = dict(type=2, add_action=True, add_reward=False, add_index=False, add_velocity=False,
              action_index="binary", piggybacking=False, add_position=False, add_positional_dist=False,
              add_positional_dist_piggy=True, add_positional_dist_type=2, add_channel_obs=False, num_bins=20)
    st.update(over)
    return st


def _eq32(got, ref64, what, t):
    ref = np.asarray(ref64, dtype=np.float64).astype(np.float32)
    bad = got != ref
    assert not bad.any(), "%s differs at slot %d in %d entries" % (what, t, int(bad.sum()))


# ---- BASELINE shapes at (or near) full size against the oracle ------------------------------------------------------

@pytest.mark.parametrize("name,E,T,kw,tables_every", [
    # configs[2] at its full batch: 4096 envs, past the 20-slot phantom phase
    ("c3_full", 4096, 40, dict(num_users=32, num_channels=20, highway_length=800), 10),
    # configs[3] shape, enough slots that `last_updated < 20` turns false for unheard vehicles
    ("c4_128x64", 64, 40, dict(num_users=128, num_channels=64, highway_length=3200), 10),
    # configs[4] at its largest vehicle count
    ("c5_256x128", 8, 30, dict(num_users=256, num_channels=128, highway_length=6400), 10),
    ("c5_64x32", 96, 40, dict(num_users=64, num_channels=32, highway_length=1600), 10),
])
def test_baseline_shapes_against_oracle(name, E, T, kw, tables_every):
    from oracle.c_oracle import COracle
    kw = dict(kw, reward_design=2, communication_range=250, mobility=True, bin_range=500, State=_state())
    seed = 4321
    orc = COracle(num_envs=E, threads=8, **kw)
    orc.reset_philox(seed)
    env = _env(E, seed=seed, **kw)
    for t in range(T):
        a = orc.philox_actions(seed, t)
        o_ref, r_ref = orc.step("my_step", a, t)
        s_ref = orc.obtain_state(o_ref, a, r_ref)
        s, r, info = env.step()                            # on-device Philox actions, fused kernel
        assert (_np(info["actions"]) == a).all()
        _eq32(_np(r), r_ref, "rews", t)
        _eq32(_np(s), s_ref, "state", t)
        _eq32(_np(info["obs"]), o_ref, "obs", t)
        if t % tables_every == tables_every - 1 or t == T - 1:
            assert (_np(env.tab_seq) == orc.tab_seq).all(), "seq table, slot %d" % t
            assert (_np(env.tab_lu) == orc.tab_lu).all(), "last_updated table, slot %d" % t
            assert (_np(env.tab_x) == orc.tab_x).all(), "xpos table, slot %d" % t
            assert (_np(env.pos_x) == orc.pos_x).all(), "pos_x, slot %d" % t
    env.close()


# ---- device topology presets (no init=) against fixtures recorded from the reference -------------------------------

@pytest.mark.parametrize("variant", ["group", "block"])
@pytest.mark.parametrize("name", ["kat1_design6_step_design", "kat2_design6_my_step", "static_design_topo_6x5"])
def test_design_topology_preset_matches_reference(name, variant):
    """enable_design_topology=True without init: Network.initialize_mobility_topology_design_test (network.py:74-79)
    as reset_kernel installs it, then the fixture's slots."""
    g, m = load_golden(name)
    kw = kwargs_from_meta(m)
    assert kw["enable_design_topology"]
    env = _env(2, variant=variant, **kw)
    assert (_np(env.pos_x)[1] == g["x0"]).all() and (_np(env.pos_y)[1] == g["y0"]).all() and (_np(env.vel)[1] == g["v0"]).all()
    for t in range(g["actions"].shape[0]):
        mode = str(g["modes"][t])
        a = np.broadcast_to(g["actions"][t], (2, env.N))
        obs, rews = env._step(mode, a, t, True)
        if m["reward_design"] in (1, 2, 5) or mode == "my_step_design":
            _eq32(_np(rews)[1], g["rews"][t], "rews", t)
        else:
            assert np.allclose(_np(rews)[1], g["rews"][t].astype(np.float32), rtol=1e-6, atol=1e-6)
        assert np.allclose(_np(env._state)[1], g["state"][t].astype(np.float32), rtol=1e-6, atol=1e-6)
        assert (_np(env.pos_x)[1] == g["pos_x"][t]).all()
        assert (_np(env.tab_seq)[1] == g["tab_seq"][t]).all() and (_np(env.tab_x)[1] == g["tab_x"][t]).all()
    env.close()


@pytest.mark.parametrize("variant", ["group", "block"])
def test_reset_mobility_env_matches_reference(variant):
    """TestEnv.reset_mobility_env (test_env.py:479-484): the fixed toy topology (network.py:81-90) with fresh tables,
    while the slot counter, last_arrival_time and the accumulators survive; then fixture kat4_toy_fixed."""
    g, m = load_golden("kat4_toy_fixed")
    env = _env(3, variant=variant, seed=11, **kwargs_from_meta(m))
    for t in range(5):                                  # scramble: random topology, tables filled, lat stamped
        env._step("my_step_ch", None, t, True)
    lat_before, t_before = env.lat.clone(), env.t
    acc_before = env._acc_count.clone()
    env.reset_mobility_env()
    assert (_np(env.pos_x)[2] == g["x0"]).all() and (_np(env.pos_y)[2] == g["y0"]).all() and (_np(env.vel)[2] == g["v0"]).all()
    assert int(env._tab_seq.abs().sum()) == 0 and int(env._tab_lu.abs().sum()) == 0 and float(env._tab_x.abs().sum()) == 0.0
    assert torch.equal(env.lat, lat_before) and env.t == t_before and torch.equal(env._acc_count, acc_before)
    for t in range(g["actions"].shape[0]):
        mode = str(g["modes"][t])
        a = np.broadcast_to(g["actions"][t], (3, env.N))
        _, rews = env._step(mode, a, t, True)
        _eq32(_np(rews)[2], g["rews"][t], "rews", t)
        assert (_np(env.pos_x)[2] == g["pos_x"][t]).all()
        assert (_np(env.tab_seq)[2] == g["tab_seq"][t]).all() and (_np(env.tab_lu)[2] == g["tab_lu"][t]).all()
        assert (_np(env.tab_x)[2] == g["tab_x"][t]).all()
        if mode != "my_step_ch":
            assert np.allclose(_np(env._state)[2], g["state"][t].astype(np.float32), rtol=1e-6, atol=1e-6)
    env.close()


# ---- compact host format == full host format == device step ---------------------------------------------------------

@pytest.mark.parametrize("n,r,E,state,extra", [
    (32, 20, 2048, _state(), {}),                                                        # the headline layout (vector rows)
    (32, 20, 1500, _state(add_channel_obs=True, add_reward=True, add_index=True, add_position=True, add_velocity=True),
     dict(enable_fingerprint=True)),                                                     # every block
    (13, 7, 1100, _state(action_index="real", num_bins=10, add_reward=True), {}),        # scalar action, odd sizes
    (6, 5, 37, _state(num_bins=37), dict(congestion_test=True, highway_length=100)),     # one chunk, toy
    (64, 8, 1024, _state(add_channel_obs=True), {}),                                     # one-CTA-per-env kernel
    (32, 20, 1024, _state(add_positional_dist_piggy=False), {}),                         # no tables at all
    (12, 4, 1024, _state(add_positional_dist_type=1), {}),                               # no compact form: falls back
])
def test_compact_host_format_is_bit_identical(n, r, E, state, extra):
    kw = dict(dict(num_users=n, num_channels=r, highway_length=25.0 * n, reward_design=2, communication_range=250,
                   mobility=True, bin_range=500, State=state), **extra)
    dev = _env(E, seed=5, **kw)
    full = _env(E, seed=5, host_format="full", **kw)
    comp = _env(E, seed=5, host_format="compact", host_threads=5, **kw)
    if state["add_positional_dist_type"] == 1:
        assert comp.host_format == "full"
    S = dev.S
    bufs = [(torch.empty((E, n, S), dtype=torch.float32).pin_memory(), torch.empty((E, n), dtype=torch.float32).pin_memory(),
             torch.empty((E, n, r), dtype=torch.float32).pin_memory()) for _ in range(2)]
    for t in range(8):
        a = dev.sample(t)
        if t == 3:
            a = a.clone(); a[0, 0] = r + 5; a[-1, -1] = -2           # out-of-range actions are clamped the same way
        ha = a.cpu().pin_memory()
        s, rw, info = dev.step(a, episode_number=t // 3, epsilon=0.5 ** t)
        full.step_host(ha, *bufs[0], episode_number=t // 3, epsilon=0.5 ** t)
        comp.lib.diral_set_option(comp._handle, b"host_nt", [-1, 0, 1][t % 3])       # every store flavour of the row assembly
        fmt = [1, 2, 3, 3, 1, 3, 2, 3][t]                # records by copy engine / zero-copy / streamed (one launch + flags)
        comp.lib.diral_set_option(comp._handle, b"host_format", fmt)
        comp.lib.diral_set_option(comp._handle, b"actions_direct", [0, 2, 1][t % 3])   # pinned actions read in place or staged
        comp.lib.diral_set_option(comp._handle, b"stream_chunks", [32, 5, 64][t % 3])
        comp.lib.diral_set_option(comp._handle, b"stream_split", int(t != 2))        # split-environment launch or the plain one
        with_obs = t == 7 or (t % 2 == 1 and fmt != 3)   # (an obs request sends format 3 down the chunked path)
        src = a.cpu().numpy().copy() if t == 5 else ha   # pageable actions: always staged
        bufs[1][0].fill_(-7.0)
        comp.step_host(src, bufs[1][0], bufs[1][1], bufs[1][2] if with_obs else None, episode_number=t // 3, epsilon=0.5 ** t)
        assert torch.equal(s.cpu(), bufs[0][0]) and torch.equal(rw.cpu(), bufs[0][1]) and torch.equal(info["obs"].cpu(), bufs[0][2])
        assert torch.equal(bufs[1][0], bufs[0][0]), "compact state rows, slot %d" % t
        assert torch.equal(bufs[1][1], bufs[0][1])
        if with_obs:
            assert torch.equal(bufs[1][2], bufs[0][2])
    assert torch.equal(dev.episode_metrics(), comp.episode_metrics())
    if state["add_positional_dist_piggy"]:
        assert torch.equal(dev._tab_seq, comp._tab_seq) and torch.equal(dev._tab_x, comp._tab_x)
    for e in (dev, full, comp):
        e.close()


def test_step_host_begin_wait_pipelines_two_env_groups():
    """diral_step_host_begin / _wait: two handles in flight at once (each its own stream and assembly threads) deliver
    the rows the synchronous call does; misuse is refused."""
    from diral_b200 import DiralError
    kw = dict(num_users=32, num_channels=20, highway_length=800, reward_design=2, communication_range=250,
              mobility=True, bin_range=500, State=_state())
    E = 1536
    sync = [_env(E, seed=s, host_threads=3, **kw) for s in (5, 6)]
    pipe = [_env(E, seed=s, host_threads=3, **kw) for s in (5, 6)]
    for e in pipe:
        e.set_host_format("compact_stream", 4, shared_pool=True)      # both groups on the process-wide assembly pool
        e.host_stream = torch.cuda.Stream(e.device)
    S = sync[0].S
    mk = lambda: (torch.empty((E, 32, S), dtype=torch.float32).pin_memory(), torch.empty((E, 32), dtype=torch.float32).pin_memory())
    ref, out = [mk(), mk()], [mk(), mk()]
    acts = [[sync[g].sample(t).cpu().pin_memory() for t in range(12)] for g in range(2)]
    torch.cuda.synchronize()
    for g in range(2):
        pipe[g].step_host_begin(acts[g][0], *out[g])
    with pytest.raises(DiralError):                     # one slot in flight per handle
        pipe[0].step_host_begin(acts[0][1], *out[0])
    with pytest.raises(DiralError):                     # ... and nothing else touches the environment meanwhile
        pipe[0].step()
    for t in range(12):
        for g in range(2):
            sync[g].step_host(acts[g][t], *ref[g])
            pipe[g].step_host_wait()
            assert torch.equal(out[g][0], ref[g][0]) and torch.equal(out[g][1], ref[g][1]), (t, g)
            out[g][0].fill_(-3.0)
            if t + 1 < 12:
                pipe[g].step_host_begin(acts[g][t + 1], *out[g])
    pipe[0].step_host_wait()                            # nothing pending: a no-op
    for g in range(2):
        assert torch.equal(sync[g]._tab_seq, pipe[g]._tab_seq) and torch.equal(sync[g].episode_metrics(), pipe[g].episode_metrics())
    big = _env(8, seed=1, **dict(kw, num_users=64, num_channels=8, highway_length=1600))
    h = (torch.empty((8, 64, big.S)).pin_memory(), torch.empty((8, 64)).pin_memory())
    with pytest.raises(DiralError):                     # not a lane-group configuration: use step_host
        big.step_host_begin(big.sample(0).cpu().pin_memory(), *h)
    for e in sync + pipe + [big]:
        e.close()


# ---- advisor findings ----------------------------------------------------------------------------------------------------

def test_update_velocity_draws_differ_between_episodes():
    """A caller following main_test.py:226-236 calls update_velocity() once per episode and never touches
    env.episode: the Philox key must move by itself."""
    from oracle.c_oracle import COracle
    kw = dict(num_users=10, num_channels=4, highway_length=400, reward_design=2, communication_range=250,
              mobility=True, mobility_vary=True, bin_range=500, State=_state())
    E, seed = 64, 5
    env = _env(E, seed=seed, **kw)
    orc = COracle(num_envs=E, **kw)
    orc.reset_philox(seed)
    deltas = []
    for episode in range(4):
        v0 = env.vel.clone()
        env.update_velocity()
        deltas.append(_np(env.vel - v0))
        orc.update_velocity(orc.philox_draws(seed, episode))
        assert (_np(env.vel) == orc.vel).all(), "episode %d" % episode
    assert env.episode == 4
    assert any((deltas[0] != d).any() for d in deltas[1:]), "every episode repeated the same draws"
    env.close()


def test_state_dict_round_trip():
    """state_dict / load_state_dict: a restored env continues bit for bit (tables, counters, lat bookkeeping)."""
    kw = dict(num_users=32, num_channels=20, highway_length=800, reward_design=3, communication_range=250,
              mobility=True, bin_range=500, enable_channel=True, State=_state())
    a_env = _env(50, seed=3, **kw)
    for t in range(30):
        a_env.step()
    snap = a_env.state_dict()
    b_env = _env(50, seed=99, **kw)
    b_env.load_state_dict(snap)
    for t in range(25):
        sa, ra, _ = a_env.step()
        sb, rb, _ = b_env.step()
        assert torch.equal(sa, sb) and torch.equal(ra, rb), t
    for name in ("_tab_seq", "_tab_lu", "_tab_x", "lat", "pos_x"):
        assert torch.equal(getattr(a_env, name), getattr(b_env, name)), name
    assert torch.equal(a_env.episode_metrics(), b_env.episode_metrics())
    assert torch.equal(a_env.network.get_information_age(a_env.t), b_env.network.get_information_age(b_env.t))
    a_env.close(); b_env.close()


def test_large_resource_count_switches_to_the_block_kernel():
    """The lane-group kernel keeps an observation row and a merge-script row per resource in shared memory; beyond
    its carve-up (R ~ 1300 at 32 vehicles) diral_create picks the one-CTA-per-env kernel instead of failing."""
    from oracle.c_oracle import COracle
    kw = dict(num_users=32, num_channels=2000, highway_length=800, reward_design=2, communication_range=250,
              mobility=True, bin_range=500, State=_state())
    E, seed = 3, 8
    env = _env(E, seed=seed, **kw)
    assert int(env.lib.diral_get_option(env._handle, b"variant")) == 2
    orc = COracle(num_envs=E, **kw)
    orc.reset_philox(seed)
    for t in range(6):
        a = orc.philox_actions(seed, t)
        o_ref, r_ref = orc.step("my_step", a, t)
        s_ref = orc.obtain_state(o_ref, a, r_ref)
        env._step("my_step", a, t, True)
        _eq32(_np(env._rews), r_ref, "rews", t)
        _eq32(_np(env._state), s_ref, "state", t)
        assert (_np(env.tab_seq) == orc.tab_seq).all()
    from diral_b200 import DiralError
    with pytest.raises(DiralError):
        _env(E, variant="group", **kw)
    env.close()


def test_table_views_need_the_piggyback_tables():
    kw = dict(num_users=6, num_channels=5, highway_length=1170, reward_design=2, communication_range=250,
              mobility=True, State=_state(add_positional_dist_piggy=False))
    env = _env(2, **kw)
    with pytest.raises(AttributeError):
        env.tab_seq
    env.close()


@pytest.mark.parametrize("n,T", [(24, 2200), (12, 4200)])
def test_partial_groups_keep_the_packed_replay(n, T):
    """N < G: lanes beyond N used to veto the 16-bit packed replay once tick > FMAX; results must (still) be exact on
    a dense highway deep into the run, where every live entry is fresh."""
    from oracle.c_oracle import COracle
    kw = dict(num_users=n, num_channels=5, highway_length=300, reward_design=2, communication_range=250,
              mobility=True, bin_range=500, State=_state())
    E, seed = 4, 17
    orc = COracle(num_envs=E, **kw)
    orc.reset_philox(seed)
    env = _env(E, seed=seed, **kw)
    for t in range(T):
        a = orc.philox_actions(seed, t)
        o_ref, r_ref = orc.step("my_step", a, t)
        env._step("my_step", a, t, True)
        if t % 200 == 199 or t > T - 5:
            _eq32(_np(env._state), orc.obtain_state(o_ref, a, r_ref), "state", t)
            assert (_np(env.tab_seq) == orc.tab_seq).all() and (_np(env.tab_lu) == orc.tab_lu).all()
            assert (_np(env.tab_x) == orc.tab_x).all()
    env.close()


# ---- split environments (tail of a batch that does not fill the device a whole number of times) ----------------------

def _set_tail_split(env, mode):
    from diral_b200._lib import check
    check(env.lib.diral_set_option(env._handle, b"tail_split", mode))


@pytest.mark.parametrize("name", ["c3_32x20_my_step", "c3_32x20_ch_d3", "c3_32x20_step_design", "c3_32x20_my_step_chanobs",
                                  "sparse_24x6_L6000_ch", "state_all_blocks_12x4"])
def test_split_environment_kernel_reproduces_fixtures(name):
    """tail_split = 2 sends EVERY environment through the 4-warps-per-environment instantiation of the lane-group
    kernel (17..32 vehicles); the fixtures recorded from the reference must come out the same."""
    from golden_util import fingerprint_args
    g, m = load_golden(name)
    if not 16 < m["num_users"] <= 32:
        pytest.skip("split environments exist for 32-lane groups only")
    E = 5
    env = _env(E, variant="group", **kwargs_from_meta(m))
    _set_tail_split(env, 2)
    env.reset(init=(g["x0"], g["y0"], g["v0"]))
    exact = m["reward_design"] in (1, 2, 5) or all(str(x) == "my_step_design" for x in g["modes"])
    for t in range(g["actions"].shape[0]):
        mode = str(g["modes"][t])
        a = np.broadcast_to(g["actions"][t], (E, env.N))
        ep, eps = fingerprint_args(m, t)
        obs, rews = env._step(mode, a, t, True, ep, eps)
        for e in (0, E - 1):
            if exact:
                _eq32(_np(rews)[e], g["rews"][t], "rews", t)
            else:
                assert np.allclose(_np(rews)[e], g["rews"][t].astype(np.float32), rtol=1e-6, atol=1e-6)
            assert np.allclose(_np(obs)[e], g["obs"][t].astype(np.float32), rtol=1e-6, atol=1e-6)
            assert np.allclose(_np(env._state)[e], g["state"][t].astype(np.float32), rtol=1e-6, atol=1e-6), t
            assert (_np(env.pos_x)[e] == g["pos_x"][t]).all()
            if m["add_positional_dist_piggy"]:
                assert (_np(env.tab_seq)[e] == g["tab_seq"][t]).all() and (_np(env.tab_lu)[e] == g["tab_lu"][t]).all()
                assert (_np(env.tab_x)[e] == g["tab_x"][t]).all()
            assert (_np(env.lat)[e] == g["lat"][t]).all(), "last_arrival_time, slot %d" % t
    env.close()


@pytest.mark.parametrize("mode,E", [("my_step", 4096), ("my_step_ch", 4096), ("my_step", 5000), ("my_step_design", 2500)])
def test_tail_split_launch_is_bit_identical(mode, E):
    """tail_split = 1 (whole waves one warp per environment + the remainder split over 4 warps each, where that remainder
    fits the device at once) and tail_split = 2 against the default single launch on the same batch: every buffer,
    table and accumulator identical."""
    kw = dict(num_users=32, num_channels=20, highway_length=800, reward_design=3 if mode == "my_step_ch" else 2,
              communication_range=250, mobility=True, bin_range=500, State=_state(add_channel_obs=(E == 5000)))
    envs = [_env(E, seed=21, **kw) for _ in range(3)]
    for k, env in enumerate(envs):
        _set_tail_split(env, k)
    for t in range(14):
        for env in envs:
            env._step(mode, None, t, True)
        for env in envs[1:]:
            for name in ("_state", "_rews", "_obs", "_tab_seq", "_tab_lu", "_tab_x", "lat", "pos_x"):
                assert torch.equal(getattr(envs[0], name), getattr(env, name)), (name, t)
    m0 = envs[0].episode_metrics()
    for env in envs[1:]:
        assert torch.equal(m0, env.episode_metrics())
    for env in envs:
        env.close()
