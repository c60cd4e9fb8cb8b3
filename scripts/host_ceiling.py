#!/usr/bin/env python
"""What bounds the end-to-end (host-buffer) number: the box's pinned device->host copy rate, the host-side row
assembly rate of the compact format by thread count, and diral_step_host itself over (host_threads, host_chunks).

    python scripts/host_ceiling.py [--gpus N] > gpurun_out/host_ceiling.json

With --gpus N the D2H ceiling is measured on N devices at once from one process (one stream per device, copies in
flight on all of them), which is what N ranks sharing the host's PCIe roots / memory controllers see.
"""
import argparse
import ctypes as C
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402

from bench import ENV_KW, E_PER_GPU, N_UE, cpu_model  # noqa: E402
from diral_b200 import TestEnv, _lib  # noqa: E402
from diral_b200.env import cfg_from_kwargs  # noqa: E402


def d2h_ceiling(n_dev, nbytes, reps=40):
    devs = [torch.device("cuda", i) for i in range(n_dev)]
    src = [torch.empty(nbytes, dtype=torch.uint8, device=d) for d in devs]
    dst = [torch.empty(nbytes, dtype=torch.uint8).pin_memory() for _ in devs]
    streams = [torch.cuda.Stream(d) for d in devs]
    for _ in range(3):
        for s, a, b in zip(streams, src, dst):
            with torch.cuda.stream(s):
                b.copy_(a, non_blocking=True)
    for d in devs:
        torch.cuda.synchronize(d)
    t0 = time.perf_counter()
    for _ in range(reps):
        for s, a, b in zip(streams, src, dst):
            with torch.cuda.stream(s):
                b.copy_(a, non_blocking=True)
    for d in devs:
        torch.cuda.synchronize(d)
    dt = time.perf_counter() - t0
    return {"devices": n_dev, "bytes": nbytes, "GBps_aggregate": n_dev * nbytes * reps / dt / 1e9,
            "us_per_copy": dt / reps * 1e6}


def expander(threads, reps=30):
    lib = _lib.load()
    cfg = cfg_from_kwargs(E_PER_GPU, 0, ENV_KW)
    A, S = E_PER_GPU * N_UE, 40
    rs = np.random.RandomState(0)
    act = rs.randint(0, 20, A).astype(np.int32)
    counts = rs.randint(0, 3, (A, 20)).astype(np.uint8)
    rews = rs.randn(A).astype(np.float32)
    out = torch.empty((A, S), dtype=torch.float32).pin_memory()
    call = lambda: lib.diral_expand_state_host(C.byref(cfg), A, act.ctypes.data, counts.ctypes.data, rews.ctypes.data, None,
                                               None, None, None, 0.0, 1.0, threads, out.data_ptr())
    for _ in range(3):
        call()
    t0 = time.perf_counter()
    for _ in range(reps):
        call()
    dt = (time.perf_counter() - t0) / reps
    return {"threads": threads, "us": dt * 1e6, "GBps_written": A * S * 4 / dt / 1e9}


def step_host(env, fmt, threads, chunks, h_act, h_state, h_rews, nt=-1, reps=40, direct=1):
    env.set_host_format(fmt, threads)
    env.lib.diral_set_option(env._handle, b"stream_chunks" if fmt == "compact_stream" else b"host_chunks", chunks)
    env.lib.diral_set_option(env._handle, b"actions_direct", direct)
    env.lib.diral_set_option(env._handle, b"tail_split", int(os.environ.get("DIRAL_TAIL_SPLIT", "0")))
    env.lib.diral_set_option(env._handle, b"host_nt", nt)
    for k in range(4):
        env.step_host(h_act[k % 8], h_state, h_rews)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for k in range(reps):
        env.step_host(h_act[k % 8], h_state, h_rews)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / reps
    tr = (C.c_double * 68)()
    n = env.lib.diral_host_trace(env._handle, tr, 68)
    return {"trace_us_last_call": [round(tr[i], 1) for i in range(n)], "format": fmt, "actions_direct": direct, "host_threads": threads, "chunks": chunks, "nt_stores": nt, "us_per_slot": dt * 1e6,
            "agent_steps_per_s": E_PER_GPU * N_UE / dt}


def pipelined(groups, threads_each, chunks, reps=60, shared=False):
    """diral_step_host_begin / _wait over `groups` handles of E / groups environments each, round robin."""
    E = E_PER_GPU // groups
    envs = [TestEnv(num_envs=E, device="cuda:0", seed=100 + g, host_threads=threads_each, **ENV_KW) for g in range(groups)]
    bufs = []
    for env in envs:
        env.set_host_format("compact_stream", threads_each, shared_pool=shared)
        env.lib.diral_set_option(env._handle, b"tail_split", int(os.environ.get("DIRAL_TAIL_SPLIT", "0")))
        env.host_stream = torch.cuda.Stream(env.device)
        env.lib.diral_set_option(env._handle, b"stream_chunks", chunks)
        acts = [env.sample(t).cpu().pin_memory() for t in range(8)]
        for t in range(40):
            env.step()
        bufs.append((acts, torch.empty((E, N_UE, env.S), dtype=torch.float32).pin_memory(),
                     torch.empty((E, N_UE), dtype=torch.float32).pin_memory()))
    torch.cuda.synchronize()

    spent = {"begin": 0.0, "wait": 0.0}

    def run(n):
        pc = time.perf_counter
        for g, env in enumerate(envs):
            env.step_host_begin(bufs[g][0][0], bufs[g][1], bufs[g][2])
        for k in range(1, n):
            for g, env in enumerate(envs):
                t0 = pc()
                env.step_host_wait()
                t1 = pc()
                env.step_host_begin(bufs[g][0][k % 8], bufs[g][1], bufs[g][2])
                spent["wait"] += t1 - t0
                spent["begin"] += pc() - t1
        for env in envs:
            env.step_host_wait()

    run(6)
    spent["begin"] = spent["wait"] = 0.0
    t0 = time.perf_counter()
    run(reps)
    dt = (time.perf_counter() - t0) / reps
    traces = []
    for env in envs:
        tr = (C.c_double * 8)()
        n = env.lib.diral_host_trace(env._handle, tr, 8)
        traces.append([round(tr[i], 1) for i in range(n)])
        env.close()
    return {"last_slot_timeline_us_per_group": traces,
            "calling_thread_us_per_slot": {k: round(v / reps * 1e6, 1) for k, v in spent.items()},"groups": groups, "envs_per_group": E, "threads_per_group": threads_each, "shared_pool": shared, "stream_chunks": chunks,
            "us_per_slot_of_all_groups": dt * 1e6, "agent_steps_per_s": E * groups * N_UE / dt}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--quick", action="store_true", help="copy ceilings and the row-assembly rate only (multi-GPU boxes are charged per GPU)")
    ap.add_argument("--pipeline", action="store_true", help="diral_step_host_begin / _wait over 2..4 env groups")
    ap.add_argument("--stream", action="store_true", help="diral_step_host only: chunked copy-engine format against the streamed one")
    args = ap.parse_args()
    cpus = len(os.sched_getaffinity(0))
    if args.pipeline:
        res = []
        total = int(os.environ.get("DIRAL_POOL_THREADS", cpus - 2))      # (the shared pool keeps the size of its first user)
        for rep in range(3):
            for groups in [int(c) for c in os.environ.get("DIRAL_PIPE_GROUPS", "1,2,3,4").split(",")]:
                for chunks in [int(c) for c in os.environ.get("DIRAL_PIPE_CHUNKS", "4,8").split(",")]:
                    res.append(pipelined(groups, total, chunks, reps=100, shared=True))
                    print(json.dumps(res[-1]), flush=True)
        return
    out = {"cpu_model": cpu_model(), "cpus_usable": cpus, "cpu_count": os.cpu_count(), "d2h": [], "expander": [], "step_host": []}
    full = E_PER_GPU * N_UE * 41 * 4
    compact = E_PER_GPU * N_UE * 24
    n = 1 if not args.stream else args.gpus + 1
    while n <= args.gpus:
        for nb in (full, compact, 256 << 20):
            out["d2h"].append(d2h_ceiling(n, nb))
        n *= 2
    for th in ([] if args.stream else sorted({1, 2, 4, 8, 12, max(cpus - 2, 1), max(cpus - 1, 1)})):
        out["expander"].append(dict(expander(th), stores="ordinary"))
        out["expander"].append(dict(expander(-th), stores="non-temporal"))
    if args.quick:
        print(json.dumps(out, indent=1))
        return
    env = TestEnv(num_envs=E_PER_GPU, device="cuda:0", seed=1234, **ENV_KW)
    acts = [env.sample(t) for t in range(8)]
    for t in range(60):
        env.step(acts[t % 8])
    h_act = [a.cpu().pin_memory() for a in acts]
    h_state = torch.empty((E_PER_GPU, N_UE, env.S), dtype=torch.float32).pin_memory()
    h_rews = torch.empty((E_PER_GPU, N_UE), dtype=torch.float32).pin_memory()
    out["step_host"].append(step_host(env, "full", 1, 4, h_act, h_state, h_rews))
    if args.stream:
        for rep in range(3):
            for th in sorted({8, max(cpus - 3, 1), max(cpus - 2, 1), max(cpus - 1, 1)}):
                out["step_host"].append(step_host(env, "compact", th, 8, h_act, h_state, h_rews, 0))
                out["step_host"].append(step_host(env, "compact_zero_copy", th, 8, h_act, h_state, h_rews, 0))
                for ch in (8, 16, 32):
                    out["step_host"].append(step_host(env, "compact_stream", th, ch, h_act, h_state, h_rews, 0))
        print(json.dumps(out, indent=1))
        return
    for th in sorted({8, 12, max(cpus - 3, 1), max(cpus - 2, 1)}):
        for ch in (2, 4, 8, 16):
            for fmt in ("compact", "compact_zero_copy"):
                out["step_host"].append(step_host(env, fmt, th, ch, h_act, h_state, h_rews, 0))
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
