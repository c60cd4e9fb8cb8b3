"""Device-resident experience memory: the reference's ``Memory`` (utils/memory.py:162-194) for a batched
GPU env, plus the learners' regrouping (algorithms/drl_drqn.py:294-377) as one gather kernel.

The reference appends ``(state, action, reward, next_state)`` tuples of per-user Python lists to a deque
and, at training time, copies ``batch_size`` windows of ``step_size`` consecutive tuples out and
re-nests them per user with four triple loops.  Here the tuples live in ring tensors on the device
(``[capacity, A, ...]`` with A = num_envs * num_users: parameter sharing makes env and user axes
exchangeable, reference README.md:9) and ``sample`` returns ``[A * batch, step, ...]`` tensors directly
in the layout ``drl_drqn.py:235-238`` reshapes to.  PyTorch only allocates; the copy is
``diral_ring_gather`` (include/diral_env.h).
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib
from ._lib import check


class Memory:
    """``Memory(max_size)`` with the reference's ``add`` / ``sample`` names (utils/memory.py:166-194)."""

    def __init__(self, max_size=1000, *, agents, state_space, device="cuda", share_next_state=True):
        self.lib = _lib.load()
        self.device = torch.device(device)
        if self.device.type != "cuda" or not torch.cuda.is_available():
            raise RuntimeError("diral_b200.replay.Memory lives on a CUDA device; there is no CPU path")
        self.capacity, self.A, self.S = int(max_size), int(agents), int(state_space)
        # main_test.py:215-217 stores next_state and then makes it the next slot's state, so by default a
        # slot's next_state is read from the following slot's state and kept only for the newest slot
        self.share_next_state = bool(share_next_state)
        f32, i32 = torch.float32, torch.int32
        self.states = torch.zeros((self.capacity, self.A, self.S), dtype=f32, device=self.device)
        self.actions = torch.zeros((self.capacity, self.A), dtype=i32, device=self.device)
        self.rewards = torch.zeros((self.capacity, self.A), dtype=f32, device=self.device)
        self.next_states = (torch.zeros((1, self.A, self.S), dtype=f32, device=self.device) if self.share_next_state
                            else torch.zeros((self.capacity, self.A, self.S), dtype=f32, device=self.device))
        self.count = 0            # experiences ever added; the deque holds the last min(count, capacity)

    def __len__(self):
        return min(self.count, self.capacity)

    def add(self, experience):
        """``memory.add((state, action, reward, next_state))`` (utils/memory.py:169-175); tensors of shape
        [E, N, S] / [E, N] (or already flattened to [A, ...]) are copied into the ring."""
        state, action, reward, next_state = experience
        slot = self.count % self.capacity
        stream = C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

        def put(ring, slot, src, dtype, shape):
            if not (isinstance(src, torch.Tensor) and src.device == self.device and src.dtype == dtype and src.is_contiguous()):
                src = torch.as_tensor(src).to(device=self.device, dtype=dtype).contiguous()
            if src.numel() != int(np.prod(shape)):
                raise ValueError("experience entry has %d elements, the ring row holds %s" % (src.numel(), shape))
            row_bytes = src.numel() * src.element_size()
            with torch.cuda.device(self.device):
                check(self.lib.diral_ring_put(ring.data_ptr(), C.c_int64(ring.shape[0]), C.c_int64(slot),
                                              C.c_int64(row_bytes), src.data_ptr(), stream))

        put(self.states, slot, state, torch.float32, (self.A, self.S))
        put(self.actions, slot, action, torch.int32, (self.A,))
        put(self.rewards, slot, reward, torch.float32, (self.A,))
        put(self.next_states, 0 if self.share_next_state else slot, next_state, torch.float32, (self.A, self.S))
        self.count += 1

    def _gather(self, ring, width, elem_bytes, start, batch, step, out):
        stream = C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)
        with torch.cuda.device(self.device):
            check(self.lib.diral_ring_gather(ring.data_ptr(), C.c_int64(ring.shape[0]), C.c_int64(self.A), C.c_int64(width),
                                             elem_bytes, start.data_ptr(), batch, step, out.data_ptr(), stream))
        return out

    def sample(self, batch_size, step_size, rng=None, idx=None):
        """utils/memory.py:177-194: ``batch_size`` windows of ``step_size`` consecutive experiences, window
        starts drawn without replacement from ``range(len - step_size)`` (``rng``: a numpy RandomState /
        Generator-like with ``choice``; ``idx`` overrides the draw).  Returns a dict of device tensors
        ``states / next_states [A*batch, step, S]``, ``actions / rewards [A*batch, step]`` with row
        ``a * batch + b`` (drl_drqn.py:294-377), and ``idx``."""
        n = len(self)
        if idx is None:
            if n - step_size < batch_size:
                raise ValueError("need at least batch_size + step_size experiences (np.random.choice would raise)")
            rng = rng if rng is not None else np.random
            idx = rng.choice(np.arange(n - step_size), size=batch_size, replace=False)
        idx = np.asarray(idx, dtype=np.int64)
        batch, step = int(len(idx)), int(step_size)
        oldest = self.count - n                      # deque index 0 = this experience
        start = torch.as_tensor((idx + oldest) % self.capacity, device=self.device)
        f32, i32, dev = torch.float32, torch.int32, self.device
        out = {"idx": idx}
        out["states"] = self._gather(self.states, self.S, 4, start, batch, step,
                                     torch.empty((self.A * batch, step, self.S), dtype=f32, device=dev))
        out["actions"] = self._gather(self.actions, 1, 4, start, batch, step,
                                      torch.empty((self.A * batch, step), dtype=i32, device=dev))
        out["rewards"] = self._gather(self.rewards, 1, 4, start, batch, step,
                                      torch.empty((self.A * batch, step), dtype=f32, device=dev))
        if self.share_next_state:
            # next_state of deque entry i is the state of entry i + 1; only the newest entry keeps its own
            nxt = self._gather(self.states, self.S, 4, (start + 1) % self.capacity, batch, step,
                               torch.empty((self.A * batch, step, self.S), dtype=f32, device=dev))
            last = np.nonzero(idx + step - 1 == n - 1)[0]
            for b in last:                           # at most one window can end at the newest experience
                nxt.view(self.A, batch, step, self.S)[:, int(b), step - 1].copy_(self.next_states[0])
            out["next_states"] = nxt
        else:
            out["next_states"] = self._gather(self.next_states, self.S, 4, start, batch, step,
                                              torch.empty((self.A * batch, step, self.S), dtype=f32, device=dev))
        return out
