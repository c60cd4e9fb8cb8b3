"""TEST INFRASTRUCTURE, not product code: CPU restatement of the reference's per-slot caller epilogue
(main_test.py:150-206 + utils/misc.py:1-12) -- information-age penalty and reward shaping.

The reference keeps this code inline in ``marl_test`` (main_test.py:14), a function that cannot run
here because the module imports TensorFlow; only ``calculate_ia_penalty`` is importable.  Parity
status: PINNED.  ``ia_penalty_sum`` against the reference's own function (tests/golden/make_golden_shaping.py);
``shape_slot`` against the reference's own loop body, sliced out of the unmodified main_test.py and executed with stub
``env`` / ``mainDRQN`` objects by tests/golden/make_golden_shaping_slots.py (fixture shaping_slots.json,
tests/test_oracle_golden.py::test_shape_slot_matches_reference_text).
"""
import numpy as np


def ia_penalty_sum(ia):
    """utils/misc.py:1-12 -- sum of (i+1) * ia[i] over the populated bins."""
    s = 0
    for i in range(len(ia)):
        if ia[i] > 0:
            s += (i + 1) * int(ia[i])
    return s


class ShapingState:
    """The loop-carried variables of main_test.py:48-56,73 for one environment."""

    def __init__(self, num_users):
        self.sum_ia_prev = 0
        self.counter = [0] * num_users
        self.previous_actions = [-1] * num_users


def shape_slot(state, ia, action, reward, num_channels, ia_averaging=False, ia_penalty_enable=False,
               ia_penalty_threshold=5, ia_penalty_value=-10, global_reward_avg=False):
    """One slot of main_test.py:150-206.  `reward` (float64 array) is modified in place like the
    reference's; returns (sum_r, collision, ia_sum)."""
    ia_sum = ia_penalty_sum(ia)                                   # :151
    ia_penalty = 0
    if ia_averaging:                                              # :153-160
        if ia_sum > state.sum_ia_prev:
            ia_penalty = -1
        elif ia_sum < state.sum_ia_prev:
            ia_penalty = 1
        state.sum_ia_prev = ia_sum
    sum_r = np.sum(reward)                                        # :174
    collision = num_channels - sum_r                              # :178
    for i in range(len(reward)):                                  # :188-206
        if ia_averaging:
            reward[i] += ia_penalty
        if ia_penalty_enable:
            if reward[i] < 1 and action[i] == state.previous_actions[i]:
                state.counter[i] += 1
            else:
                state.counter[i] = 0
            if state.counter[i] > ia_penalty_threshold:
                reward[i] = ia_penalty_value
            state.previous_actions[i] = action[i]
        if global_reward_avg:
            reward[i] = reward[i] + sum_r / len(reward)
    return sum_r, collision, ia_sum
