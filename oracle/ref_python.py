"""Times the UNMODIFIED reference ``TestEnv`` (oracle/_ref/envs, placed there by oracle/install_ref.py) on host cores.

Test / bench infrastructure only -- nothing under diral_b200/ imports this.  The reference is imported with the two
shims SURVEY.md 8(c) lists and no source edits: ``sys.path`` += envs/ (py2 implicit-relative imports,
envs/test_env.py:4, envs/network.py:9) and a stub ``matplotlib`` (imported at network.py:6 for the dead ``plot_fc``).

Timed region = what one slot of the GPU path replaces: ``my_step`` (envs/test_env.py:124) or ``my_step_ch`` (:351)
followed by ``obtain_state`` (:527), per BASELINE.md section 3: one core / one env, and P processes each stepping its
own env for a fixed wall-clock budget (the reference is single-threaded and GIL-bound, so processes not threads).
"""
from __future__ import annotations

import contextlib
import io
import multiprocessing as mp
import os
import random
import sys
import time
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF_ENVS = os.path.join(HERE, "_ref", "envs")


def available() -> bool:
    return all(os.path.exists(os.path.join(REF_ENVS, f)) for f in ("vehicle.py", "network.py", "test_env.py"))


def load_reference():
    if not available():
        raise RuntimeError("oracle/_ref/envs is empty: run `python oracle/install_ref.py` where /root/reference exists")
    if "matplotlib" not in sys.modules:
        mpl = types.ModuleType("matplotlib")
        plt = types.ModuleType("matplotlib.pyplot")
        mpl.pyplot = plt
        sys.modules["matplotlib"] = mpl
        sys.modules["matplotlib.pyplot"] = plt
    if REF_ENVS not in sys.path:
        sys.path.insert(0, REF_ENVS)
    import test_env  # noqa: E402  (the reference module)
    return test_env.TestEnv


def cpu_model() -> str:
    try:
        with open("/proc/cpuinfo") as f:
            for line in f:
                if line.startswith("model name"):
                    return line.split(":", 1)[1].strip()
    except OSError:
        pass
    return "unknown"


def _make_env(env_kw, seed):
    TestEnv = load_reference()
    np.random.seed(seed); random.seed(seed)
    with contextlib.redirect_stdout(io.StringIO()):
        return TestEnv(**env_kw)


def _slots(env, mode, seconds, warm):
    """Step one reference env for `seconds` of wall clock after `warm` slots; returns (slots, elapsed)."""
    step = getattr(env, mode)
    n, t = env.NUM_USERS, 0
    sink = io.StringIO()
    with contextlib.redirect_stdout(sink):
        for _ in range(warm):
            a = env.sample(); obs, rews = step(a, t); env.obtain_state(obs, a, rews); t += 1
        done = 0
        t0 = time.perf_counter()
        while True:
            a = env.sample(); obs, rews = step(a, t); env.obtain_state(obs, a, rews); t += 1
            done += 1
            el = time.perf_counter() - t0
            if el >= seconds:
                return done, el


def _worker(args):
    env_kw, mode, seconds, warm, seed = args
    env = _make_env(env_kw, seed)
    done, el = _slots(env, mode, seconds, warm)
    return done, el, env.NUM_USERS


def time_reference(env_kw, mode="my_step", seconds=3.0, warm=25, processes=None, seed=1234):
    """{"one_core": agent-steps/s of one env on one core, "all_cores": aggregate of P processes, "P", "cpu_model"}."""
    done, el, n = _worker((env_kw, mode, seconds, warm, seed))
    one = done * n / el
    P = processes or (len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1))
    ctx = mp.get_context("fork")
    with ctx.Pool(P) as pool:
        res = pool.map(_worker, [(env_kw, mode, seconds, warm, seed + 1 + k) for k in range(P)])
    allc = sum(d * nn / e for d, e, nn in res)
    return {"one_core": one, "all_cores_P_processes": allc, "P": P, "cpu_model": cpu_model(),
            "seconds_per_leg": seconds, "slots_one_core": done,
            "what": "unmodified reference TestEnv.%s + obtain_state (envs/test_env.py), 1 env per process" % mode}


if __name__ == "__main__":
    # `python -m oracle.ref_python '<json kwargs>' [mode] [seconds] [processes]` -> one JSON line.  bench.py runs it
    # as a child process: forking workers from a process that holds a CUDA context is not safe.
    import json
    kw = json.loads(sys.argv[1])
    mode = sys.argv[2] if len(sys.argv) > 2 else "my_step"
    seconds = float(sys.argv[3]) if len(sys.argv) > 3 else 3.0
    procs = int(sys.argv[4]) if len(sys.argv) > 4 and int(sys.argv[4]) > 0 else None
    print(json.dumps(time_reference(kw, mode=mode, seconds=seconds, processes=procs)))
