"""CPU, world_size 2 over gloo: the data-parallel plumbing (bucketed all-reduce of the flat gradient buffer driven by
the backward's progress callbacks, volume sharding) without any GPU."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from brats21_b200 import parallel


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        n = 1000
        flat = torch.arange(n, dtype=torch.float32) * (rank + 1)
        red = parallel.BucketReducer(flat, bucket_elems=128)
        red.start()
        for upto in (0, 100, 128, 500, 777):  # progress in uneven steps; buckets fire only when complete
            red.progress(upto)
            assert red.next == upto // 128
        red.finish()
        assert red.next == len(red.bounds) == 8 and red.launched == 8
        expect = torch.arange(n, dtype=torch.float32) * sum(range(1, world + 1))
        ok = torch.equal(flat, expect)
        # a second backward re-uses the reducer
        flat.copy_(torch.ones(n) * (rank + 1))
        red.start()
        red.finish()
        ok = ok and torch.equal(flat, torch.full((n,), float(sum(range(1, world + 1)))))
        out[rank] = bool(ok)
    finally:
        dist.destroy_process_group()


def test_bucket_reducer_world2_gloo():
    world, port = 2, _free_port()
    with mp.Manager() as m:
        out = m.dict()
        mp.spawn(_worker, args=(world, port, out), nprocs=world, join=True)
        assert dict(out) == {0: True, 1: True}


def test_shard_indices_partition():
    for n, world in ((219, 8), (5, 8), (16, 4), (1, 1)):
        parts = [parallel.shard_indices(n, r, world) for r in range(world)]
        assert sorted(i for p in parts for i in p) == list(range(n))
        assert max(len(p) for p in parts) - min(len(p) for p in parts) <= 1
    assert parallel.shard_indices(219, 0, 8)[:3] == [0, 8, 16] and len(parallel.shard_indices(219, 7, 8)) == 27
