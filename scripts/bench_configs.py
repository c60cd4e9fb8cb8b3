#!/usr/bin/env python
"""Device-resident throughput of every BASELINE.json config (supplementary to bench.py, which is the
contract for configs[2]).  Prints one JSON line per config: slot time, agent-steps/s, roofline fraction
against the algorithmic bytes of SURVEY.md 8(d).  L2 is flushed before every timed slot."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from diral_b200 import TestEnv  # noqa: E402

STATE = dict(type=2, add_action=True, add_reward=False, add_index=False, add_velocity=False,
             action_index="binary", piggybacking=False, add_position=False, add_positional_dist=False,
             add_positional_dist_piggy=True, add_positional_dist_type=2, add_channel_obs=False, num_bins=20)
CONFIGS = [
    ("C1 4x3 toy", 4096, dict(num_users=4, num_channels=3, highway_length=100, congestion_test=True), "my_step"),
    ("C2 6x5", 1024, dict(num_users=6, num_channels=5, highway_length=1170), "my_step"),
    ("C3 32x20", 4096, dict(num_users=32, num_channels=20, highway_length=800), "my_step"),
    ("C3 32x20 PRR", 4096, dict(num_users=32, num_channels=20, highway_length=800, reward_design=3), "my_step_ch"),
    ("C4 128x64 (1/8 of 16384)", 2048, dict(num_users=128, num_channels=64, highway_length=3200), "my_step"),
    # configs[4]: vehicle-count sweep at 8192 envs, L = 25 N, R = max(3, N / 2)
    ("C5 4x3", 8192, dict(num_users=4, num_channels=3, highway_length=100), "my_step"),
    ("C5 8x4", 8192, dict(num_users=8, num_channels=4, highway_length=200), "my_step"),
    ("C5 16x8", 8192, dict(num_users=16, num_channels=8, highway_length=400), "my_step"),
    ("C5 32x16", 8192, dict(num_users=32, num_channels=16, highway_length=800), "my_step"),
    ("C5 64x32", 8192, dict(num_users=64, num_channels=32, highway_length=1600), "my_step"),
    ("C5 128x64", 8192, dict(num_users=128, num_channels=64, highway_length=3200), "my_step"),
    ("C5 256x128", 8192, dict(num_users=256, num_channels=128, highway_length=6400), "my_step"),
]


def main():
    peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(
        os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0
    flush_w = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    flush_r = torch.ones(64 << 20, dtype=torch.float32, device="cuda")
    only = sys.argv[1:] and sys.argv[1]
    for name, E, kw, mode in CONFIGS:
        if only and only not in name:
            continue
        kw = dict(dict(reward_design=2, communication_range=250, mobility=True, bin_range=500, State=STATE), **kw)
        env = TestEnv(num_envs=E, device="cuda", seed=1, variant=os.environ.get("DIRAL_VARIANT", "auto"), **kw)
        for t in range(30):
            env._step(mode, None, t, True)
        steps = 50
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
        for k in range(steps):
            flush_w.zero_(); flush_r.sum()
            ev[k][0].record(); env._step(mode, None, 30 + k, True); ev[k][1].record()
        torch.cuda.synchronize()
        ms = sum(a.elapsed_time(b) for a, b in ev) / steps
        n, r, b = env.N, env.R, env.B
        alg = (32 * n * n + n * (36 + 8 * r + 4 * b) + (8 * n * n if mode == "my_step_ch" else 0)) * E
        row = {"config": name, "kernel": env.kernel, "envs": E, "mode": mode, "us_per_slot": ms * 1e3,
               "agent_steps_per_s": E * n / (ms / 1e3), "algorithmic_GBps": alg / (ms / 1e3) / 1e9,
               "roofline_frac": alg / (ms / 1e3) / 1e9 / peak}
        if n <= 32:
            # fused rollout: T slots of every environment in ONE launch (diral_rollout), on-device actions; the
            # state of these small configurations stays in L1/L2, so this is a latency number, not an HBM one
            T = 200
            env.rollout(T, mode)
            r0, r1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            r0.record(); env.rollout(T, mode); r1.record()
            torch.cuda.synchronize()
            us = r0.elapsed_time(r1) * 1e3 / T
            row["rollout_us_per_slot"] = us
            row["rollout_agent_steps_per_s"] = E * n / (us / 1e6)
        print(json.dumps(row))
        env.close(); del env
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
