"""The RealNeS side of the reference, batched on the device (last row).

* ``get_neighbor_dist`` / ``get_neighbor_dist2``: the view-based positional distribution that
  ``RealnessEnv`` builds from a received neighbour table (reference envs/realness_env.py:52-85 / :87-118), for a
  batch of tables in the wire layout of ``MA_NeighborTableEntry`` (envs/ma_messages_pb2.py: pos_x, pos_y float32;
  seq_num, last_update int32) -- 16 bytes per entry, which is exactly how the tensor is laid out, so a table
  parsed off the socket can be copied to the device without repacking.
* ``SemiPersistentScheduling``: the 3GPP SPS baseline (reference algorithms/v2x_sps.py) as one state machine per
  agent, with the reference's constructor arguments and ``step(selection_window)``.

PyTorch only allocates; the work is ``diral_wire_vpd`` / ``diral_sps_step`` (include/diral_env.h).  There is
no CPU path.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib
from ._lib import DiralSpsCfg, check

WIRE_DTYPE = np.dtype([("pos_x", "<f4"), ("pos_y", "<f4"), ("seq_num", "<i4"), ("last_update", "<i4")])


def _cuda(device):
    device = torch.device(device)
    if device.type != "cuda" or not torch.cuda.is_available():
        raise RuntimeError("diral_b200.realness runs on a CUDA device; there is no CPU path")
    return device if device.index is not None else torch.device("cuda", torch.cuda.current_device())


def pack_tables(pos_x, pos_y, seq_num, last_update, device="cuda"):
    """[M, N] arrays -> one [M, N, 4] int32 device tensor holding MA_NeighborTableEntry records bit for bit."""
    device = _cuda(device)
    rec = np.empty(np.shape(pos_x), dtype=WIRE_DTYPE)
    rec["pos_x"], rec["pos_y"], rec["seq_num"], rec["last_update"] = pos_x, pos_y, seq_num, last_update
    return torch.from_numpy(rec.view(np.int32).reshape(rec.shape + (4,))).to(device)


def _vpd(tables, observer, pos_dist, state_bins, state_range, age_limit):
    lib = _lib.load()
    if tables.dtype != torch.int32 or tables.dim() != 3 or tables.shape[2] != 4 or not tables.is_contiguous():
        raise ValueError("tables must be a contiguous int32 tensor [M, N, 4] (see pack_tables)")
    device = _cuda(tables.device)
    M, N = int(tables.shape[0]), int(tables.shape[1])
    observer = torch.as_tensor(observer, dtype=torch.int32, device=device).contiguous()
    if observer.shape != (M,):
        raise ValueError("observer must hold one 0-based user id per table")
    if M and (int(observer.min()) < 0 or int(observer.max()) >= N):
        raise ValueError("observer ids must be in [0, N)")
    out = torch.empty((M, int(state_bins)), dtype=torch.float32, device=device)
    stream = C.c_void_p(torch.cuda.current_stream(device).cuda_stream)
    with torch.cuda.device(device):
        check(lib.diral_wire_vpd(tables.data_ptr(), observer.data_ptr(), C.c_int64(M), C.c_int32(N), C.c_int32(pos_dist),
                                 C.c_int32(int(state_bins)), C.c_double(float(state_range)), C.c_int32(int(age_limit)),
                                 out.data_ptr(), stream))
    return out


def get_neighbor_dist(tables, observer, state_bins=10, age_limit=20):
    """``RealnessEnv.get_neighbor_dist(tx_id, pos_of_neighbors)`` (realness_env.py:52-85) for M tables."""
    return _vpd(tables, observer, 1, state_bins, 1.0, age_limit)


def get_neighbor_dist2(tables, observer, state_bins=10, state_range=250, age_limit=20):
    """``RealnessEnv.get_neighbor_dist2(tx_id, pos_of_neighbors)`` (realness_env.py:87-118) for M tables."""
    return _vpd(tables, observer, 2, state_bins, state_range, age_limit)


class SemiPersistentScheduling:
    """``SemiPersistentScheduling(user, selection_window, RSSI_threshold)`` (v2x_sps.py:8-22) for ``agents``
    users at once.  ``step(selection_window[, draws])`` takes a float64 device tensor [agents, window_len] of
    averaged RSSI values and returns the int32 subframe chosen by every agent (v2x_sps.py:76-104)."""

    def __init__(self, agents, selection_window, RSSI_threshold, *, device="cuda", seed=0, init=None):
        self.lib = _lib.load()
        self.device = _cuda(device)
        self.A = int(agents)
        self.selection_window_size = int(selection_window)
        self.RSSI_threshold = float(RSSI_threshold)
        self.inc_dB, self.prob_resource_keep = 3.0, 0.8                    # v2x_sps.py:19,22
        self.seed, self.t = int(seed), 0
        if init is None:   # txSubframe = randint(0, window), reselection_counter = randint(5, 15)  (v2x_sps.py:14-16)
            g = torch.Generator(device="cpu").manual_seed(self.seed)
            tx = torch.randint(0, self.selection_window_size + 1, (self.A,), generator=g, dtype=torch.int32)
            cnt = torch.randint(5, 16, (self.A,), generator=g, dtype=torch.int32)
        else:
            tx, cnt = (torch.as_tensor(np.asarray(v), dtype=torch.int32) for v in init)
        self.prev_action = tx.to(self.device).contiguous()
        self.reselection_counter = cnt.to(self.device).contiguous()
        self.flags = torch.zeros((self.A,), dtype=torch.int32, device=self.device)
        self._actions = torch.empty((self.A,), dtype=torch.int32, device=self.device)

    def step(self, selection_window, draws=None):
        w = selection_window
        if w.dtype != torch.float64 or w.dim() != 2 or w.shape[0] != self.A or not w.is_contiguous() or w.device != self.device:
            raise ValueError("selection_window must be a contiguous float64 tensor [agents, window_len] on %s" % self.device)
        if draws is not None and (draws.dtype != torch.float64 or tuple(draws.shape) != (self.A, 3) or not draws.is_contiguous()):
            raise ValueError("draws must be a contiguous float64 tensor [agents, 3]")
        wn = int(w.shape[1])
        cfg = DiralSpsCfg(self.RSSI_threshold, self.inc_dB, self.prob_resource_keep, wn / 5)   # v2x_sps.py:40
        stream = C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)
        with torch.cuda.device(self.device):
            check(self.lib.diral_sps_step(C.c_int64(self.A), C.c_int32(wn), w.data_ptr(), C.byref(cfg),
                                          draws.data_ptr() if draws is not None else None, C.c_uint64(self.seed),
                                          C.c_int64(self.t), self.prev_action.data_ptr(), self.reselection_counter.data_ptr(),
                                          self._actions.data_ptr(), self.flags.data_ptr(), stream))
        self.t += 1
        return self._actions
