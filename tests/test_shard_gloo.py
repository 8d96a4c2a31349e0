"""N>1 host logic on CPU: world_size-2 gloo processes exercise the sharding, the index-blob broadcast
and the counter reduction that bench.py / a multi-GPU caller use (NCCL on the GPU box)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from bgreat_b200 import shard
    # 1. shards tile the read set exactly
    n = 1000003
    lo, hi = shard.shard_range(n, rank, world)
    sizes = [torch.zeros(1, dtype=torch.int64) for _ in range(world)]
    dist.all_gather(sizes, torch.tensor([hi - lo], dtype=torch.int64))
    assert sum(int(x) for x in sizes) == n
    assert max(int(x) for x in sizes) - min(int(x) for x in sizes) <= 1
    # 2. the blob arrives bit-identical on every rank
    rng = np.random.default_rng(7)
    ref = torch.from_numpy(rng.integers(0, 256, size=123457, dtype=np.uint8))
    got = shard.broadcast_blob(ref.clone() if rank == 0 else None, src=0, device=torch.device("cpu"))
    assert torch.equal(got, ref)
    # 3. counters add up
    c = shard.reduce_counters({"reads": hi - lo, "aligned": rank + 1, "no_overlap": 10 * rank})
    assert c == {"reads": n, "aligned": sum(range(1, world + 1)), "no_overlap": 10 * sum(range(world))}
    open(os.path.join(out_dir, "ok%d" % rank), "w").close()
    dist.destroy_process_group()


def test_world_size_2_gloo(tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    assert all(os.path.exists(tmp_path / ("ok%d" % r)) for r in range(world))


def test_shard_range_properties():
    from bgreat_b200 import shard
    for n in (0, 1, 7, 10_000_000):
        for world in (1, 2, 3, 8):
            spans = [shard.shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
