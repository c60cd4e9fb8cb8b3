"""TEST INFRASTRUCTURE, not product code: CPU restatement of the reference's experience memory
(utils/memory.py:162-194, ``Memory``) and of the learners' per-user regrouping loops
(algorithms/drl_drqn.py:294-377).

Parity status: ``Memory`` is pinned against the reference's own class by
tests/golden/make_golden_replay.py (fixture tests/golden/replay_memory.npz); the regrouping functions
live in a module that imports TensorFlow, so ``regroup`` is pinned against their text sliced out of the unmodified
file and executed by tests/golden/make_golden_regroup.py (fixture replay_regroup.npz).
"""
from collections import deque

import numpy as np


class Memory:
    def __init__(self, max_size=1000):
        self.buffer = deque(maxlen=max_size)                      # utils/memory.py:166-167

    def add(self, experience):
        self.buffer.append(experience)                            # :169-175

    def sample(self, batch_size, step_size, rng=np.random):
        idx = rng.choice(np.arange(len(self.buffer) - step_size), size=batch_size, replace=False)   # :184-185
        res = []
        for i in idx:                                             # :189-193
            temp_buffer = []
            for j in range(step_size):
                temp_buffer.append(self.buffer[i + j])
            res.append(temp_buffer)
        return res, idx


def regroup(batch, field, num_users):
    """drl_drqn.py:294-377 (get_states_user / get_actions_user / get_rewards_user / get_next_states_user):
    batch[b][k] = (states, actions, rewards, next_states) -> array [user][b][k][...]."""
    out = []
    for user in range(num_users):
        per_user = []
        for each in batch:
            per_batch = []
            for step_i in each:
                per_batch.append(step_i[field][user])
            per_user.append(per_batch)
        out.append(per_user)
    return np.array(out)
