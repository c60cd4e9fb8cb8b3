#!/usr/bin/env python
"""Pins oracle/replay.py::regroup to the reference's own text (run in the build container: needs /root/reference).

``get_states_user / get_actions_user / get_rewards_user / get_next_states_user`` (algorithms/drl_drqn.py:294-377)
are methods of a class whose module imports TensorFlow.  The four methods themselves need only NumPy, so this script
slices their text out of the UNMODIFIED file (from ``def get_states_user`` up to the next ``def sample``), dedents
it, executes the ``def`` statements in a namespace with ``np`` / ``sys``, and calls them with a stub ``self`` that has
``num_users``.  Inputs and outputs go to tests/golden/replay_regroup.npz; nothing is copied into the repo.
"""
import os
import sys
import textwrap
import types

import numpy as np

REF = os.environ.get("DIRAL_REFERENCE", "/root/reference")


def sliced_methods():
    lines = open(os.path.join(REF, "algorithms", "drl_drqn.py")).read().split("\n")
    i0 = next(i for i, l in enumerate(lines) if l.strip().startswith("def get_states_user("))
    i1 = next(i for i, l in enumerate(lines) if l.strip().startswith("def sample(") and i > i0)
    ns = {"np": np, "sys": sys}
    exec(compile(textwrap.dedent("\n".join(lines[i0:i1])), "drl_drqn.py[%d:%d]" % (i0 + 1, i1), "exec"), ns)
    return ns, (i0 + 1, i1)


def main():
    ns, span = sliced_methods()
    rs = np.random.RandomState(21)
    U, S, BATCH, STEP = 6, 5, 4, 3
    me = types.SimpleNamespace(num_users=U)
    # batch[b][k] = (states [U][S], actions [U], rewards [U], next_states [U][S]) as Memory.sample returns it
    st = rs.rand(BATCH, STEP, U, S); ac = rs.randint(0, 5, size=(BATCH, STEP, U)); rw = rs.randn(BATCH, STEP, U)
    nx = rs.rand(BATCH, STEP, U, S)
    batch = [[(st[b, k], ac[b, k], rw[b, k], nx[b, k]) for k in range(STEP)] for b in range(BATCH)]
    out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "replay_regroup.npz")
    np.savez_compressed(out, states=st, actions=ac, rewards=rw, next_states=nx,
                        out_states=ns["get_states_user"](me, batch), out_actions=ns["get_actions_user"](me, batch),
                        out_rewards=ns["get_rewards_user"](me, batch), out_next_states=ns["get_next_states_user"](me, batch),
                        span=np.array(span))
    print(out, "sliced drl_drqn.py lines %d-%d" % span)


if __name__ == "__main__":
    main()
