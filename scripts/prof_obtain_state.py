#!/usr/bin/env python
"""Steps the two un-fused State variants at C3 (VPD type 1, sorted direct distribution) for an ncu capture of
obtain_state_kernel:  ncu --set full -k regex:obtain_state -s 20 -c 2 -o gpurun_out/prof_obtain python scripts/prof_obtain_state.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from bench import STATE  # noqa: E402
from diral_b200 import TestEnv  # noqa: E402

which = sys.argv[1] if len(sys.argv) > 1 else "vpd1"
state = dict(STATE, add_positional_dist_type=1) if which == "vpd1" else dict(STATE, add_positional_dist=True, add_positional_dist_piggy=False)
env = TestEnv(num_envs=4096, device="cuda", seed=1, num_users=32, num_channels=20, highway_length=800, reward_design=2,
              communication_range=250, mobility=True, bin_range=500, State=state)
for t in range(40):
    env.step()
torch.cuda.synchronize()
