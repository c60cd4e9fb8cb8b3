#!/usr/bin/env python
"""BASELINE configs[3]: 128 UE x 64 resources, 16384 envs sharded across the GPUs of one box (2048 envs per GPU at 8).
Launch with torchrun (one rank per GPU); prints one JSON line from rank 0: device time per slot (CUDA events, L2
flushed before every timed slot, max over ranks) and whole-job agent-steps/s."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402
from diral_b200 import TestEnv  # noqa: E402
from scripts.bench_configs import STATE  # noqa: E402


def main():
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    total = 16384
    E = total // world
    kw = dict(num_users=128, num_channels=64, highway_length=3200, reward_design=2, communication_range=250, mobility=True,
              bin_range=500, State=STATE)
    env = TestEnv(num_envs=E, device="cuda:%d" % local, seed=1, env_offset=rank * E, **kw)
    flush_w = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    flush_r = torch.ones(64 << 20, dtype=torch.float32, device="cuda")
    for t in range(30):
        env._step("my_step", None, t, True)
    steps = 40
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    if hasattr(torch.cuda, "_sleep"):
        torch.cuda._sleep(int(2.0e7))              # let the host run ahead of the device (see bench.py)
    for k in range(steps):
        flush_w.zero_(); flush_r.sum()
        ev[k][0].record(); env._step("my_step", None, 30 + k, True); ev[k][1].record()
    torch.cuda.synchronize()
    ms = torch.tensor([sum(a.elapsed_time(b) for a, b in ev) / steps], dtype=torch.float64, device="cuda")
    allms = [torch.zeros_like(ms) for _ in range(world)]
    if world > 1:
        dist.all_gather(allms, ms)
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    else:
        allms = [ms.clone()]
    if rank == 0:
        t = float(ms.item())
        alg = (32 * 128 * 128 + 128 * (36 + 8 * 64 + 4 * 20)) * E
        print(json.dumps({"config": "configs[3]: 128 UE x 64 res, 16384 envs over %d GPU(s)" % world, "n_gpus": world,
                          "envs_per_gpu": E, "us_per_slot": t * 1e3, "agent_steps_per_s": total * 128 / (t / 1e3),
                          "roofline_frac_per_gpu": alg / (t / 1e3) / 1e9 / 6545.9,
                          "per_rank_us": [float(x.item()) * 1e3 for x in allms]}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
