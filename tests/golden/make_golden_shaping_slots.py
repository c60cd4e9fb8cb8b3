#!/usr/bin/env python
"""Pins oracle/shaping.py::shape_slot to the reference's own text (run in the build container: needs /root/reference).

The per-slot caller epilogue lives inline in ``marl_test`` (main_test.py:14), whose module imports TensorFlow and
cannot be imported here.  The loop body itself needs nothing but NumPy, so this script slices the statements from
``ia = env.network.get_information_age(time_step)`` up to (not including) ``log_reward_slot.append(sum_r)``
(main_test.py:150-206) out of the UNMODIFIED file, dedents them, and executes them slot after slot in a namespace that
supplies what the surrounding function would: a stub ``env`` (recorded information-age vectors, a no-op
``obtain_state``), a stub ``mainDRQN``, the reference's own ``calculate_ia_penalty`` (utils/misc.py), the option
flags and the loop-carried variables initialised as main_test.py:48-56,73 does.  Nothing is copied into the repo:
the text is executed, the inputs and outputs are recorded to tests/golden/shaping_slots.json.
"""
import json
import os
import sys
import textwrap

import numpy as np

REF = os.environ.get("DIRAL_REFERENCE", "/root/reference")
sys.path.insert(0, REF)
from utils.misc import calculate_ia_penalty  # noqa: E402  (the reference function)

START = "ia = env.network.get_information_age(time_step)"
END = "log_reward_slot.append(sum_r)"


def slot_body():
    lines = open(os.path.join(REF, "main_test.py")).read().split("\n")
    i0 = next(i for i, l in enumerate(lines) if l.strip() == START)
    i1 = next(i for i, l in enumerate(lines) if l.strip() == END and i > i0)
    return compile(textwrap.dedent("\n".join(lines[i0:i1])), "main_test.py[%d:%d]" % (i0 + 1, i1), "exec"), (i0 + 1, i1)


class _Net:
    def __init__(self):
        self.ia = None

    def get_information_age(self, t):
        return self.ia


class _Env:
    def __init__(self):
        self.network = _Net()

    def obtain_state(self, *a):
        return None


class _Drqn:
    def get_eps(self):
        return 1.0


def main():
    code, span = slot_body()
    rs = np.random.RandomState(12)
    cases = []
    for opts in (dict(), dict(global_reward_avg=True), dict(ia_averaging=True), dict(ia_averaging=True, global_reward_avg=True),
                 dict(ia_penalty_enable=True, ia_penalty_threshold=2, ia_penalty_value=-10),
                 dict(ia_penalty_enable=True, ia_penalty_threshold=0, ia_penalty_value=-3.5, global_reward_avg=True),
                 dict(ia_averaging=True, ia_penalty_enable=True, ia_penalty_threshold=1, ia_penalty_value=-7.5, global_reward_avg=True)):
        for n, R in ((4, 3), (12, 3), (32, 20)):
            full = dict(ia_averaging=False, ia_penalty_enable=False, ia_penalty_threshold=5, ia_penalty_value=-10,
                        global_reward_avg=False)
            full.update(opts)
            env = _Env()
            ns = dict(env=env, mainDRQN=_Drqn(), calculate_ia_penalty=calculate_ia_penalty, np=np, num_channels=R,
                      log_ia_slot=[], cum_r=[0], cum_r_slots=[0], cum_collision=[0], cum_collision_slots=[0], episode=0,
                      sum_ia_prev=0, ia_penalty_counter=np.zeros(n), previous_actions=np.zeros(n) - 1,   # main_test.py:55,56,73
                      obs=None, **full)
            slots = []
            prev = rs.randint(0, R, size=n)
            for t in range(16):
                action = np.where(rs.rand(n) < 0.75, prev, rs.randint(0, R, size=n)).astype(np.int32)
                prev = action
                reward = rs.choice([1.0, 0.0, -2.0, -3.0, -0.5, 0.25], size=n)
                ia = (rs.randint(0, 25, size=100) * (rs.rand(100) < 0.2)).astype(np.int64)
                env.network.ia = ia
                ns.update(action=action.copy(), reward=reward.copy(), time_step=t)
                exec(code, ns)
                slots.append(dict(ia=ia.tolist(), action=action.tolist(), reward_in=reward.tolist(),
                                  reward_out=[float(x) for x in ns["reward"]], sum_r=float(ns["sum_r"]),
                                  collision=float(ns["collision"]), ia_sum=int(ns["ia_sum"]),
                                  sum_ia_prev=int(ns["sum_ia_prev"]),
                                  counter=[int(x) for x in ns["ia_penalty_counter"]],
                                  previous_actions=[int(x) for x in ns["previous_actions"]]))
            cases.append(dict(num_users=n, num_channels=R, opts=full, slots=slots))
    out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "shaping_slots.json")
    json.dump(dict(source="main_test.py lines %d-%d, executed" % span, cases=cases), open(out, "w"))
    print(out, len(cases), "cases; sliced main_test.py lines %d-%d" % span)


if __name__ == "__main__":
    main()
