#!/usr/bin/env python
"""Headline benchmark: agent-steps/sec of the V2V environment hot path (BASELINE.json metric).

One "step" = one time slot over the whole env batch: actions in -> (obs, rewards, state) out, i.e.
one my_step + one obtain_state of the reference (envs/test_env.py:124,527), here ONE fused kernel
launch.  Workload at every N: BASELINE configs[2] per GPU -- 4096 envs x 32 UE x 20 resources,
shipped State block, B=20, W=500, C=250, L=800, reward design 2 -- so N GPUs step N*4096 envs
(weak scaling; the env batch shards with no data-path collective, the only exchange is the
110-element episode-metric all-reduce every 25 slots).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

Timing: per-step CUDA events on the launching stream, L2 flushed (untimed) before every timed step so
the tables really come from HBM, summed over the K steps, max over ranks.  `e2e` drives the C-ABI
host-buffer entry point (pinned host actions in, state + rewards out) and is wall-clock.
`--impl reference` times the CPU port of the reference algorithm (oracle/, plain C, all host threads)
on a bounded sample of the same workload; the Python reference itself cannot travel to the GPU box.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

print_line = print
METRIC = "agent-steps/sec (4096 envs x 32 UE x 20 res per GPU)"
UNIT = "agent-steps/s"
E_PER_GPU, N_UE, N_RES, N_BINS = 4096, 32, 20, 20
EPISODE = 25     # main_test.py:226 episode_interval

STATE = dict(type=2, add_action=True, add_reward=False, add_index=False, add_velocity=False,
             action_index="binary", piggybacking=False, add_position=False, add_positional_dist=False,
             add_positional_dist_piggy=True, add_positional_dist_type=2, add_channel_obs=False, num_bins=N_BINS)
ENV_KW = dict(num_users=N_UE, num_channels=N_RES, highway_length=800, reward_design=2, communication_range=250,
              mobility=True, bin_range=500, State=STATE)


def algorithmic_bytes_per_env_step(n, r, b):
    """table read + write at 16 B/entry each way, plus the O(N) vectors."""
    return 32 * n * n + n * (36 + 8 * r + 4 * b)


def workload_config(n_gpus):
    return {"workload": "configs[2]: 4096 envs x 32 UE x 20 resources per GPU, my_step + obtain_state (fused)",
            "envs_per_gpu": E_PER_GPU, "num_users": N_UE, "num_channels": N_RES, "num_bins": N_BINS,
            "highway_length": 800, "communication_range": 250, "bin_range": 500, "reward_design": 2,
            "mode": "my_step", "state": "one-hot action + VPD type 2 (S=40)", "parallelism": "env-batch dp%d" % n_gpus,
            "l2": "flushed (256 MiB write + 256 MiB read, untimed) before every timed step"}


# ----------------------------------------------------------------------------- CPU arm
def cpu_port_run(steps, warmup, target_seconds=20.0):
    """Time the C restatement of the reference (oracle/) on all host cores, bounded sample."""
    from oracle.c_oracle import COracle
    cores = os.cpu_count() or 1
    # calibrate the sample size: one slot of 64 envs per thread
    probe_E = 64 * cores
    orc = COracle(num_envs=probe_E, threads=cores, **ENV_KW)
    orc.reset_philox(1234)
    acts = orc.philox_actions(1234, 0)
    t0 = time.perf_counter()
    for t in range(3):
        o, r = orc.step("my_step", acts, t); orc.obtain_state(o, acts, r)
    per_env_slot = (time.perf_counter() - t0) / (3 * probe_E)
    E = int(target_seconds / max(per_env_slot * (steps + warmup), 1e-9))
    E = max(cores, min(E_PER_GPU, (E // cores) * cores))
    orc = COracle(num_envs=E, threads=cores, **ENV_KW)
    orc.reset_philox(1234)
    actions = [orc.philox_actions(1234, t) for t in range(steps + warmup)]
    for t in range(warmup):
        o, r = orc.step("my_step", actions[t], t); orc.obtain_state(o, actions[t], r)
    t0 = time.perf_counter()
    for t in range(warmup, warmup + steps):
        o, r = orc.step("my_step", actions[t], t); orc.obtain_state(o, actions[t], r)
    dt = time.perf_counter() - t0
    value = E * N_UE * steps / dt
    sample = "%d envs x %d UE x %d res, %d slots after %d warm-up, %d pthreads" % (E, N_UE, N_RES, steps, warmup, cores)
    return value, dt, cores, sample


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps, warmup = args.steps, args.warmup
    value, dt, cores, sample = cpu_port_run(steps, warmup, target_seconds=60.0)
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
            "warmup": warmup, "ms_per_step": dt / steps * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": workload_config(args.gpus),
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0,
            "note": "CPU port (oracle/diral_oracle.c) of the reference's my_step + obtain_state; the reference is "
                    "pure Python and cannot travel to the GPU box (it ran ~60x slower per core than this port in "
                    "the build container, SURVEY.md section 6)"}
    print_line(json.dumps(line))


# ----------------------------------------------------------------------------- clocks
class ClockSampler:
    """Samples SM clock / throttle reasons through NVML while the timed region runs."""

    def __init__(self, index):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._thread = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception as exc:  # noqa: BLE001
            self.nv = None
            self.err = str(exc)

    def _loop(self):
        nv = self.nv
        names = {getattr(nv, k): k for k in dir(nv) if k.startswith("nvmlClocksEventReason") or k.startswith("nvmlClocksThrottleReason")}
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    mask = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:  # noqa: BLE001
                    mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in names.items():
                    if isinstance(bit, int) and bit and (mask & bit) == bit and bin(bit).count("1") == 1:
                        self.reasons.add(name.replace("nvmlClocksEventReason", "").replace("nvmlClocksThrottleReason", ""))
            except Exception:  # noqa: BLE001
                pass
            time.sleep(0.002)

    def start(self):
        if self.nv:
            self._thread = threading.Thread(target=self._loop, daemon=True)
            self._thread.start()

    def stop(self):
        if self._thread:
            self._stop.set(); self._thread.join()
        benign = {"GpuIdle", "None", "ApplicationsClocksSetting"}
        return {"sm_mhz": statistics.median(self.samples) if self.samples else None, "sm_max_mhz": self.max_mhz,
                "reasons": sorted(r for r in self.reasons if r not in benign), "samples": len(self.samples)}


# ----------------------------------------------------------------------------- GPU arm
def run_gpu(args):
    import torch
    import torch.distributed as dist
    from diral_b200 import TestEnv
    from diral_b200.dist import all_reduce_metrics, init_from_env

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback")
    rank, world, local = init_from_env("nccl")
    if world != args.gpus and world > 1:
        raise SystemExit("--gpus %d but WORLD_SIZE=%d" % (args.gpus, world))
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    steps, warmup = args.steps, max(args.warmup, 3)

    env = TestEnv(num_envs=E_PER_GPU, device=dev, seed=1234, env_offset=rank * E_PER_GPU, **ENV_KW)
    stream = torch.cuda.current_stream(dev)
    flush_w = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    flush_r = torch.ones(64 << 20, dtype=torch.float32, device=dev)     # 256 MiB
    side = torch.cuda.Stream(dev)

    def flush():
        flush_w.zero_()            # evicts (and writes back) whatever the last step left dirty
        flush_r.sum()              # then fill L2 with clean lines so the timed step pays no write-back for them

    def barrier():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()

    # pre-generated device-resident actions (inputs resident in HBM when the timed region starts)
    n_act = 64
    actions = [env.sample(t) for t in range(n_act)]
    for t in range(100):                          # leave the 20-slot phantom phase
        env.step(actions[t % n_act])
    for _ in range(warmup):
        flush(); env.step(actions[0])
    ev0 = [torch.cuda.Event(enable_timing=True) for _ in range(steps)]
    ev1 = [torch.cuda.Event(enable_timing=True) for _ in range(steps)]
    pending = []
    clocks = ClockSampler(local)
    launches0 = env.launch_count()
    barrier()
    clocks.start()
    # a short spin kernel lets the host run ahead of the device, so that no timed interval contains the
    # host's own launch latency (with 8 ranks per box the host loop is the slower one at first)
    if hasattr(torch.cuda, "_sleep"):
        torch.cuda._sleep(int(2.0e7))
    t_host0 = time.perf_counter()
    for k in range(steps):
        flush()
        ev0[k].record(stream)
        env.step(actions[k % n_act])
        if (k + 1) % EPISODE == 0:                # end of episode: metric vector + the one collective
            vec = env.episode_metrics()
            if world > 1:
                side.wait_stream(stream)
                with torch.cuda.stream(side):
                    out = vec.clone()
                    pending.append((out, all_reduce_metrics(out, async_op=True)))
        ev1[k].record(stream)
    host_us_per_step = (time.perf_counter() - t_host0) / steps * 1e6
    for _, work in pending:
        if work is not None:
            work.wait()
    barrier()
    clk = clocks.stop()
    gpu_launches = env.launch_count() - launches0
    ms = sum(a.elapsed_time(b) for a, b in zip(ev0, ev1))
    t_ms = torch.tensor([ms], dtype=torch.float64, device=dev)
    per_rank = None
    if world > 1:
        # every rank's own device time and median SM clock travel to rank 0: the headline uses the MAX
        mine = torch.tensor([ms / steps, float(clk.get("sm_mhz") or 0.0), float(len(clk.get("reasons") or []))],
                            dtype=torch.float64, device=dev)
        allr = [torch.zeros_like(mine) for _ in range(world)]
        dist.all_gather(allr, mine)
        per_rank = {"ms_per_step": [float(t[0]) for t in allr], "sm_mhz": [float(t[1]) for t in allr],
                    "throttle_reasons": [int(t[2]) for t in allr]}
        dist.all_reduce(t_ms, op=dist.ReduceOp.MAX)
    ms = float(t_ms.item())
    value = world * E_PER_GPU * N_UE * steps / (ms / 1e3)

    # --- steady state without the flush (state stays in the 126 MB L2): supplementary, not the headline
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for k in range(steps):
        env.step(actions[k % n_act])
    e1.record(stream)
    barrier()
    warm_value = E_PER_GPU * N_UE * steps / (e0.elapsed_time(e1) / 1e3)

    # --- e2e through the C ABI with host buffers (pinned), copies inside the timed region
    S = env.S
    h_act = [a.cpu().pin_memory() for a in actions[:8]]
    h_state = torch.empty((E_PER_GPU, N_UE, S), dtype=torch.float32).pin_memory()
    h_rews = torch.empty((E_PER_GPU, N_UE), dtype=torch.float32).pin_memory()
    e2e_steps = max(10, min(steps, 200))
    for k in range(3):
        env.step_host(h_act[k % 8], h_state, h_rews)
    barrier()
    t0 = time.perf_counter()
    for k in range(e2e_steps):
        env.step_host(h_act[k % 8], h_state, h_rews)
    torch.cuda.synchronize(dev)
    e2e_s = time.perf_counter() - t0
    t_e2e = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t_e2e, op=dist.ReduceOp.MAX)
    e2e_value = world * E_PER_GPU * N_UE * e2e_steps / float(t_e2e.item())
    h2d = E_PER_GPU * N_UE * 4
    d2h = E_PER_GPU * N_UE * (S + 1) * 4

    if rank != 0:
        return
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak, peak_src = json.load(open(peaks_path))["hbm_gbs"], "measured (MEASURED_PEAKS.json hbm_gbs)"
    else:
        peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
    alg = algorithmic_bytes_per_env_step(N_UE, N_RES, N_BINS) * E_PER_GPU
    k_ms = ms / steps                             # the fused slot kernel is the only kernel of a step
    achieved = alg / (k_ms / 1e3) / 1e9
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tpath):
        traffic = json.load(open(tpath)).get("step_group_kernel_dram_bytes_per_launch")
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": steps, "warmup": warmup,
            "ms_per_step": k_ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic", "config": workload_config(world), "clocks": clk,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "steps": e2e_steps, "api": "diral_step_host (C ABI, pinned host buffers, synchronous)"},
            "gpu_launches": gpu_launches,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "kernel": "step_group_kernel<32>", "algorithmic_bytes_per_launch": alg,
                         "peak_source": peak_src},
            "extra": {"value_l2_resident_no_flush": warm_value,
                      "note": "value_l2_resident_no_flush = same loop without the L2 flush (state fits the 126 MB L2)"}}
    line["extra"]["host_enqueue_us_per_step"] = host_us_per_step
    if per_rank is not None:
        line["extra"]["per_rank"] = per_rank
    if world == 1 and not args.no_cpu:
        v, dt, cores, sample = cpu_port_run(20, 3, target_seconds=15.0)
        line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample}
    print_line(json.dumps(line))


def _claim_stdout():
    """Libraries (NCCL's version banner, for one) print to fd 1; the driver wants exactly one JSON line
    there.  Point fd 1 at stderr for the run and hand back a writer for the real stdout."""
    sys.stdout.flush()
    real = os.dup(1)
    os.dup2(2, 1)
    return lambda text: os.write(real, (text + "\n").encode())


def main():
    emit = _claim_stdout()
    global print_line
    print_line = emit
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=500)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)
        try:
            import torch.distributed as dist
            if dist.is_initialized():
                dist.destroy_process_group()
        except Exception:  # noqa: BLE001
            pass


if __name__ == "__main__":
    main()
