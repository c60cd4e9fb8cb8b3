"""Helpers shared by the parity tests: load a golden fixture (tests/golden/*.npz, produced by
running the unmodified reference, see make_golden.py) and rebuild the reference-style kwargs."""
import glob
import os

import numpy as np

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


AUX_FIXTURES = {"replay_memory", "replay_regroup", "wire_vpd", "sps"}      # fixtures of the "next" rows, not env rollouts


def golden_names():
    names = sorted(os.path.splitext(os.path.basename(p))[0] for p in glob.glob(os.path.join(GOLDEN_DIR, "*.npz")))
    return [n for n in names if n not in AUX_FIXTURES]


def load_golden(name):
    g = dict(np.load(os.path.join(GOLDEN_DIR, name + ".npz")))
    meta = {k[5:]: g[k].item() for k in list(g) if k.startswith("meta_")}
    return g, meta


def kwargs_from_meta(m):
    """EnvironmentTest-style kwargs (reference test_env.py:12-48) from the fixture's metadata."""
    state = dict(type=m["state_type"], add_action=m["add_action"], add_reward=m["add_reward"],
                 add_index=m["add_index"], add_velocity=m["add_velocity"],
                 action_index="binary" if m["action_binary"] else "real", piggybacking=False,
                 add_position=m["add_position"], add_positional_dist=m["add_positional_dist"],
                 add_positional_dist_piggy=m["add_positional_dist_piggy"],
                 add_positional_dist_type=m["add_positional_dist_type"],
                 add_channel_obs=m["add_channel_obs"], num_bins=m["num_bins"])
    return dict(num_users=m["num_users"], num_channels=m["num_channels"], mobility=m["mobility"],
                mobility_vary=m["mobility_vary"], enable_design_topology=m["enable_design_topology"],
                highway_length=m["highway_length"], enable_fingerprint=m["enable_fingerprint"],
                reward_design=m["reward_design"], communication_range=m["communication_range"],
                bin_range=m["bin_range"], congestion_test=m["congestion_test"], State=state)


def fingerprint_args(m, t):
    return (t // 25, 0.9992 ** (t // 25)) if m["fingerprint_args"] else (0, 1)
