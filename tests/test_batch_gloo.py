"""The multi-GPU batch path on CPU: world_size 2 over gloo.  Sharding and the result gather are the product's
(mincostflow_b200/batch.py); the per-instance solve is stood in by the CPU oracle, since the engine has no CPU path."""
import os
import socket
import sys

import numpy as np
import pytest
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _records(ids):
    from mincostflow_b200 import batch, instances
    from oracle import oracle
    recs = []
    for i in ids:
        p = instances.netgen8(8, seed=13502460 + i)
        r, flow, pi, _, _ = oracle.solve(p, config=oracle.default_config())
        recs.append(batch.make_record(i, r.status, r.iterations, r.total_cost, flow, pi))
    return recs


def _worker(rank, world, port, count, out_path):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    from mincostflow_b200 import batch
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    mine = batch.shard(count, world, rank)
    got = batch.gather_records(_records(mine), count, dist=dist)
    if rank == 0:
        np.save(out_path, got)
    dist.barrier()
    dist.destroy_process_group()


def test_shard_is_a_partition():
    from mincostflow_b200 import batch
    for count in (0, 1, 5, 64):
        for world in (1, 2, 4, 8):
            ids = sorted(i for r in range(world) for i in batch.shard(count, world, r))
            assert ids == list(range(count))


def test_checksum_is_order_sensitive():
    from mincostflow_b200 import batch
    a = np.array([1, 2, 3, 4], np.int64)
    assert batch.checksum(a) != batch.checksum(a[::-1].copy()) and batch.checksum(a) == batch.checksum(a.copy())


@pytest.mark.timeout(300)
def test_two_rank_gather_matches_single_process(tmp_path):
    from mincostflow_b200 import batch
    count = 5                                                      # odd: ranks hold 3 and 2 instances (ragged gather)
    out = str(tmp_path / "gathered.npy")
    mp.spawn(_worker, args=(2, _free_port(), count, out), nprocs=2, join=True)
    got = np.load(out)
    want = batch.gather_records(_records(range(count)), count)
    assert got.shape == (count, len(batch.RECORD_FIELDS))
    assert np.array_equal(got, want)
    assert (got[:, 1] == 1).all()                                  # all optimal
