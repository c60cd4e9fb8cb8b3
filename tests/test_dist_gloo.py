"""Multi-process host logic on CPU (gloo, world_size 2): env-range sharding and the one collective of
the system -- the all-reduce(sum) of the 110-element end-of-episode metric vector."""
import os
import socket
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_range_partitions_exactly():
    from diral_b200.dist import shard_range
    for total in (1, 7, 4096, 16384, 8191):
        for world in (1, 2, 3, 8):
            spans = [shard_range(total, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_range(8, 2, 2)


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


def _worker(rank, world, port, total_envs, out):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    from diral_b200.dist import all_reduce_metrics, init_from_env, metrics_dict, shard_range
    r, w, _ = init_from_env("gloo")
    lo, hi = shard_range(total_envs, r, w)
    # a metric vector whose global sum is known in closed form: every env contributes its index
    vec = torch.zeros(110, dtype=torch.float64)
    idx = torch.arange(lo, hi, dtype=torch.float64)
    vec[0] = idx.sum(); vec[2] = (hi - lo) * 3; vec[3] = (hi - lo) * 4; vec[4] = (hi - lo) * 32
    vec[10 + r] = 1
    work = all_reduce_metrics(vec, async_op=True)
    work.wait()
    d = metrics_dict(vec)
    ok = (vec[0].item() == total_envs * (total_envs - 1) / 2 and d["agent_steps"] == total_envs * 32
          and abs(d["prr"] - 0.75) < 1e-12 and d["information_age"][:world] == [1.0] * world)
    out[rank] = bool(ok)
    dist.destroy_process_group()


def test_metric_all_reduce_world2():
    world, total = 2, 4097
    mgr = mp.Manager()
    out = mgr.dict()
    port = _free_port()
    mp.spawn(_worker, args=(world, port, total, out), nprocs=world, join=True)
    assert dict(out) == {0: True, 1: True}
