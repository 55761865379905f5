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


K = 8                                                               # NETGEN-8 2^8: n = 256, m = 2048


def _solve(i):
    from mincostflow_b200 import instances
    from oracle import oracle
    p = instances.netgen8(K, seed=13502460 + i)
    r, flow, pi, _, _ = oracle.solve(p, config=oracle.default_config())
    return p, r, flow, pi


def _buffer(ids, per, width):
    """This rank's gather buffer: one full record {status, pivots, cost, flow[m], pi[n]} per instance (CPU tensors under gloo)."""
    import torch
    from mincostflow_b200 import batch
    buf = batch.new_buffer(per, width)
    for row, i in enumerate(ids):
        p, r, flow, pi = _solve(i)
        batch.pack_record(buf[row], i, r.status, r.iterations, r.total_cost, torch.from_numpy(flow), torch.from_numpy(pi))
    return buf


def _worker(rank, world, port, count, out_path):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    from mincostflow_b200 import batch
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    mine = batch.shard(count, world, rank)
    per = (count + world - 1) // world
    got = batch.gather_records(_buffer(mine, per, batch.record_width(2048, 256)), count, dist=dist)
    if rank == 0:
        assert all(v["ok"] for v in got.values())                  # checksums re-computed from the arrays that arrived
        np.savez(out_path, **{f"{k}_{i}": (np.asarray(v[k]) if k in ("flow", "pi") else np.int64(v[k])) for i, v in got.items()
                              for k in ("status", "pivots", "total_cost", "flow", "pi")})
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


def test_torch_checksum_equals_numpy_checksum():
    import torch
    from mincostflow_b200 import batch
    rng = np.random.default_rng(1)
    a = rng.integers(-2**62, 2**62, 1000, dtype=np.int64)
    assert batch.checksum_t(torch.from_numpy(a)) == batch.checksum(a)
    assert batch.checksum_t(torch.empty(0, dtype=torch.int64)) == batch.checksum(np.empty(0, np.int64)) == 0


@pytest.mark.timeout(300)
def test_two_rank_gather_delivers_every_flow_and_potential(tmp_path):
    """rank 0 ends up with {status, pivots, total_cost, flow[m], pi[n]} of EVERY instance, bit-identical to what a
    single process computes."""
    count = 5                                                      # odd: ranks hold 3 and 2 instances (ragged gather)
    out = str(tmp_path / "gathered.npz")
    mp.spawn(_worker, args=(2, _free_port(), count, out), nprocs=2, join=True)
    got = np.load(out)
    for i in range(count):
        p, r, flow, pi = _solve(i)
        assert int(got[f"status_{i}"]) == r.status == 1
        assert int(got[f"pivots_{i}"]) == r.iterations and int(got[f"total_cost_{i}"]) == r.total_cost
        assert np.array_equal(got[f"flow_{i}"], flow) and np.array_equal(got[f"pi_{i}"], pi)


def test_gather_flags_a_corrupted_record():
    import torch
    from mincostflow_b200 import batch
    buf = _buffer([0], 1, batch.record_width(2048, 256))
    assert batch.gather_records(buf, 1)[0]["ok"]
    buf[0, batch.HEADER + 7] += 1                                  # one flow value changes on the way
    assert not batch.gather_records(buf, 1)[0]["ok"]
