"""Pin the CPU oracle (oracle/diral_oracle.c) to the reference: replay every golden fixture
(outputs of the unmodified reference env) and demand bit-identical float64 results."""
import numpy as np
import pytest

from golden_util import fingerprint_args, golden_names, kwargs_from_meta, load_golden
from oracle.c_oracle import COracle


def _same(a, b, what, t):
    a = np.asarray(a, dtype=np.float64); b = np.asarray(b, dtype=np.float64)
    ok = (a == b) | (np.isnan(a) & np.isnan(b))
    assert ok.all(), "%s differs at slot %d: %d entries, max |d|=%g" % (
        what, t, (~ok).sum(), np.nanmax(np.abs(a - b)))
    # also the sign of zero (the reference produces -0.0 rewards under my_step_ch design 2)
    assert (np.signbit(a) == np.signbit(b))[~np.isnan(a)].all(), "%s sign-of-zero differs at slot %d" % (what, t)


@pytest.mark.parametrize("name", golden_names())
def test_c_oracle_matches_reference(name):
    g, m = load_golden(name)
    orc = COracle(num_envs=1, **kwargs_from_meta(m))
    assert orc.S == m["state_space"]
    if "trace" in g:
        orc.load_trace(g["trace"])
    orc.reset(g["x0"][None], g["y0"][None], g["v0"][None])
    T = g["actions"].shape[0]
    for t in range(T):
        mode = str(g["modes"][t])
        obs, rews = orc.step(mode, g["actions"][t][None], t)
        ep, eps = fingerprint_args(m, t)
        state = orc.obtain_state(obs, g["actions"][t][None], rews, ep, eps)
        _same(obs[0], g["obs"][t], "obs", t)
        _same(rews[0], g["rews"][t], "rews", t)
        _same(orc.pos_x[0], g["pos_x"][t], "pos_x", t)
        _same(state[0], g["state"][t], "state", t)
        assert (orc.tab_seq[0] == g["tab_seq"][t]).all(), "seq table, slot %d" % t
        assert (orc.tab_lu[0] == g["tab_lu"][t]).all(), "last_updated table, slot %d" % t
        _same(orc.tab_x[0], g["tab_x"][t], "xpos table", t)
        _same(orc.tab_y[0], g["tab_y"][t], "ypos table", t)
        assert (orc.lat[0] == g["lat"][t]).all(), "last_arrival_time, slot %d" % t
        assert (orc.information_age(t)[0] == g["ia"][t]).all(), "information age, slot %d" % t
        if m["episode_interval"] and t % m["episode_interval"] == m["episode_interval"] - 1:
            orc.update_velocity(g["draws"][t][None])
        _same(orc.vel[0], g["vel"][t], "velocity", t)


def test_survey_known_answers():
    """SURVEY.md section 2b KAT-1..4, read back from the fixtures (guards the fixtures themselves)."""
    g, _ = load_golden("kat1_design6_step_design")
    assert g["rews"].tolist() == [[1] * 6, [-2, -2, 1, 1, 1, 1], [-3, -3, -3, 1, 1, 1]]
    vpd0 = g["state"][0][0][5:]
    assert {i: round(v, 6) for i, v in enumerate(vpd0) if v} == {9: 0.8, 13: 0.2}
    g, _ = load_golden("kat2_design6_my_step")
    assert g["rews"].tolist() == [[0, 1, 1, 1, 1, 0], [-2, -2, 1, 1, 1, 1], [0, 1, 0, 1, 1, 1]]
    assert g["obs"][0][0].tolist() == [0, 195, 1e5, 1e5, 1e5]
    g, _ = load_golden("kat3_design6_ch_d2")
    assert g["rews"][2].tolist() == [-0.0, 1, -0.5, 1, 1, 1] and np.signbit(g["rews"][2][0])
    g, _ = load_golden("kat4_toy_fixed")
    assert g["rews"][:3].tolist() == [[-2, 1, 1, -2], [0, 1, 1, 0], [-4] * 4]
    assert g["pos_x"][4].tolist() == [5.5, 10.0, 9.25, 12.5]


def test_ia_penalty_sum_matches_reference():
    """oracle/shaping.py::ia_penalty_sum against outputs of the reference's utils/misc.py function."""
    import json
    import os
    from golden_util import GOLDEN_DIR
    from oracle.shaping import ia_penalty_sum
    cases = json.load(open(os.path.join(GOLDEN_DIR, "shaping_ia_sums.json")))
    assert len(cases) > 60
    for c in cases:
        assert ia_penalty_sum(c["ia"]) == c["sum"]


def test_replay_memory_matches_reference():
    """oracle/replay.py::Memory against windows sampled by the reference's utils/memory.py::Memory."""
    import os
    from golden_util import GOLDEN_DIR
    from oracle.replay import Memory, regroup
    g = np.load(os.path.join(GOLDEN_DIR, "replay_memory.npz"))
    N, S, CAP, T, BATCH, STEP = (int(v) for v in g["shape"])
    mem = Memory(max_size=CAP)
    k = 0
    for t in range(T):
        mem.add((g["states"][t], g["actions"][t], g["rewards"][t], g["states"][t + 1]))
        if k < len(g["at"]) and t == int(g["at"][k]):
            rs = np.random.RandomState(100 + t)
            batch, idx = mem.sample(BATCH, STEP, rng=rs)
            assert (idx == g["idx"][k]).all()
            flat = np.array([[np.concatenate([e[0].ravel(), e[1].ravel(), e[2].ravel(), e[3].ravel()]) for e in w]
                             for w in batch])
            assert (flat == g["samples"][k]).all()
            st = regroup(batch, 0, N)
            assert st.shape == (N, BATCH, STEP, S) and (st[2, 1, 0] == batch[1][0][0][2]).all()
            k += 1
    assert k == len(g["at"]) and k > 3


def test_wire_vpd_oracle_matches_reference():
    """oracle/wire.py against RealnessEnv.get_neighbor_dist / get_neighbor_dist2 run unmodified
    (tests/golden/make_golden_wire.py), bit for bit in float64."""
    import os
    from oracle import wire
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "wire_vpd.npz"))
    for i in range(int(g["ncases"])):
        x, y, lu, obs, o1, o2 = (g["c%d_%s" % (i, k)] for k in ("x", "y", "lu", "obs", "o1", "o2"))
        bins, rng = int(g["c%d_bins" % i]), float(g["c%d_rng" % i])
        for m in range(x.shape[0]):
            r2 = wire.neighbor_dist2(int(obs[m]), x[m], y[m], lu[m], bins, rng)
            assert (np.asarray(r2, dtype=np.float64) == o2[m]).all(), (i, m)
            if m > 0:            # table 0 of every case has a zero norm: NaNs in the reference, not compared
                with np.errstate(all="ignore"):
                    r1 = wire.neighbor_dist(int(obs[m]), x[m], y[m], lu[m], bins)
                assert np.array_equal(np.asarray(r1, dtype=np.float64), o1[m], equal_nan=True), (i, m)


def test_sps_oracle_matches_reference():
    """oracle/wire.py::SpsBank against SemiPersistentScheduling (algorithms/v2x_sps.py) stepped with scripted draws."""
    import os
    from oracle import wire
    s = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "sps.npz"))
    for i in range(int(s["ncases"])):
        bank = wire.SpsBank(s["c%d_tx0" % i], s["c%d_c0" % i], float(s["c%d_thr" % i]))
        W, D = s["c%d_windows" % i], s["c%d_draws" % i]
        for t in range(W.shape[0]):
            a, f = bank.step(W[t], D[t])
            assert (a == s["c%d_acts" % i][t]).all(), (i, t)
            assert (bank.prev == s["c%d_prev" % i][t]).all() and (bank.counter == s["c%d_cnt" % i][t]).all(), (i, t)
            assert not f.any()


def test_shape_slot_matches_reference_text():
    """oracle/shaping.py::shape_slot against the reference's own loop body -- main_test.py:150-206, sliced out of the
    unmodified file and executed by tests/golden/make_golden_shaping_slots.py: shaped rewards, slot sums and every
    loop-carried variable, float64 bit for bit, over all option combinations."""
    import json
    import os
    from golden_util import GOLDEN_DIR
    from oracle.shaping import ShapingState, shape_slot
    fx = json.load(open(os.path.join(GOLDEN_DIR, "shaping_slots.json")))
    assert len(fx["cases"]) >= 20
    for c in fx["cases"]:
        st = ShapingState(c["num_users"])
        for k, s in enumerate(c["slots"]):
            reward = np.array(s["reward_in"], dtype=np.float64)
            sum_r, coll, ia_sum = shape_slot(st, s["ia"], s["action"], reward, c["num_channels"], **c["opts"])
            assert reward.tolist() == s["reward_out"], (c["opts"], k)
            assert (sum_r, coll, ia_sum) == (s["sum_r"], s["collision"], s["ia_sum"]), (c["opts"], k)
            if c["opts"]["ia_averaging"]:
                assert st.sum_ia_prev == s["sum_ia_prev"]
            if c["opts"]["ia_penalty_enable"]:
                assert [int(x) for x in st.counter] == s["counter"]
                assert [int(x) for x in st.previous_actions] == s["previous_actions"]


def test_regroup_matches_reference_text():
    """oracle/replay.py::regroup against get_states_user / get_actions_user / get_rewards_user / get_next_states_user
    (algorithms/drl_drqn.py:294-377), sliced out of the unmodified file and executed by make_golden_regroup.py."""
    import os
    from golden_util import GOLDEN_DIR
    from oracle.replay import regroup
    g = np.load(os.path.join(GOLDEN_DIR, "replay_regroup.npz"))
    st, ac, rw, nx = g["states"], g["actions"], g["rewards"], g["next_states"]
    BATCH, STEP, U = ac.shape
    batch = [[(st[b, k], ac[b, k], rw[b, k], nx[b, k]) for k in range(STEP)] for b in range(BATCH)]
    for field, name in enumerate(("states", "actions", "rewards", "next_states")):
        got = regroup(batch, field, U)
        want = g["out_" + name]
        assert got.shape == want.shape and got.dtype == want.dtype and (got == want).all(), name
