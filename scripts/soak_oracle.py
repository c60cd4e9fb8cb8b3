#!/usr/bin/env python
"""Long-horizon parity soak: BASELINE shapes stepped for thousands of slots against the C oracle (well past the 2 047
slots after which the 16-bit packed table keys of the lane-group kernel are relative to a moving base), rewards / obs /
state rows every slot, tables every 100.  python scripts/soak_oracle.py [slots]"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402
from bench import STATE  # noqa: E402
from diral_b200 import TestEnv  # noqa: E402
from oracle.c_oracle import COracle  # noqa: E402

slots = int(sys.argv[1]) if len(sys.argv) > 1 else 5000
CASES = [("C3 32x20", 24, dict(num_users=32, num_channels=20, highway_length=800), "my_step"),
         ("sparse 32x8, L=4000", 24, dict(num_users=32, num_channels=8, highway_length=4000), "my_step"),
         ("C3 32x20 PRR (design 3)", 24, dict(num_users=32, num_channels=20, highway_length=800, reward_design=3, enable_channel=True), "my_step_ch"),
         ("24x10 design mode", 24, dict(num_users=24, num_channels=10, highway_length=600), "my_step_design"),
         ("64x32 (pair kernel)", 6, dict(num_users=64, num_channels=32, highway_length=1600), "my_step"),
         ("48x12 PRR (pair kernel)", 6, dict(num_users=48, num_channels=12, highway_length=1200, reward_design=4, enable_channel=True), "my_step_ch"),
         ("128x64 (row kernel)", 3, dict(num_users=128, num_channels=64, highway_length=3200), "my_step"),
         ("80x24 (round-1 block kernel)", 3, dict(num_users=80, num_channels=24, highway_length=2000), "my_step")]
np32 = lambda t: t.detach().cpu().numpy()
bad = 0
for name, E, kw, mode in CASES:
    kw = dict(dict(reward_design=2, communication_range=250, mobility=True, bin_range=500, State=STATE), **kw)
    n = slots if kw["num_users"] <= 32 else max(slots // 4, 600)
    orc = COracle(num_envs=E, threads=8, **kw)
    orc.reset_philox(99)
    env = TestEnv(num_envs=E, device="cuda:0", seed=99, **kw)
    t0 = time.time()
    for t in range(n):
        a = orc.philox_actions(99, t)
        o_ref, r_ref = orc.step(mode, a, t)
        s_ref = orc.obtain_state(o_ref, a, r_ref)
        env._step(mode, None, t, True)
        s, r, info = env._state, env._rews, {"obs": env._obs}
        ok = (np32(env._actions) == a).all() and (np32(r) == r_ref.astype(np.float32)).all() and (np32(s) == s_ref.astype(np.float32)).all() \
            and (np32(info["obs"]) == o_ref.astype(np.float32)).all()
        if t % 100 == 99 or t == n - 1:
            ok = ok and (np32(env.tab_seq) == orc.tab_seq).all() and (np32(env.tab_lu) == orc.tab_lu).all() \
                and (np32(env.tab_x) == orc.tab_x).all() and (np32(env.pos_x) == orc.pos_x).all()
        if not ok:
            bad += 1
            print("MISMATCH", name, "slot", t, flush=True)
            break
    print("%s: %d slots x %d envs, kernel %s, %s, %.0f s" % (name, n, E, env.kernel, "ok" if ok else "FAILED", time.time() - t0), flush=True)
    env.close()
sys.exit(1 if bad else 0)
