"""Batches of independent instances over several GPUs (BASELINE.json config 5).

A single solve is sequential across pivots and stays on one GPU; a batch shards by instance.  One process per GPU
(`torch.distributed`, NCCL over NVLink; `gloo` in the CPU tests): rank r solves instances r, r + W, r + 2W, ...
back to back and the fixed-size result records are gathered on rank 0.  No data-path collective exists."""
from __future__ import annotations

import numpy as np

RECORD_FIELDS = ("instance", "status", "pivots", "total_cost", "flow_checksum", "pi_checksum")


def shard(count: int, world: int, rank: int) -> list[int]:
    """Instance ids owned by `rank` (round robin, like mcf_solve_batch's device assignment)."""
    return list(range(rank, count, world))


def checksum(a: np.ndarray) -> int:
    """Order-sensitive 63-bit checksum of an int64 array (so that permuted results do not collide)."""
    a = np.ascontiguousarray(a, np.int64).view(np.uint64)
    with np.errstate(over="ignore"):
        w = (np.arange(1, a.size + 1, dtype=np.uint64) * np.uint64(0x9E3779B97F4A7C15)) | np.uint64(1)
        return int((a * w).sum(dtype=np.uint64) & np.uint64(0x7FFFFFFFFFFFFFFF))


def make_record(instance: int, status: int, pivots: int, total_cost: int, flows: np.ndarray, pis: np.ndarray) -> list[int]:
    return [int(instance), int(status), int(pivots), int(total_cost), checksum(flows), checksum(pis)]


def gather_records(local: list[list[int]], count: int, dist=None, device=None):
    """All ranks call; rank 0 gets an int64 array [count, len(RECORD_FIELDS)] ordered by instance id, others None."""
    import torch
    world = dist.get_world_size() if dist is not None else 1
    rank = dist.get_rank() if dist is not None else 0
    per = (count + world - 1) // world
    buf = torch.full((per, len(RECORD_FIELDS)), -1, dtype=torch.int64, device=device)
    if local:
        buf[:len(local)] = torch.tensor(local, dtype=torch.int64, device=device)
    if dist is None:
        got = [buf]
    else:
        got = [torch.empty_like(buf) for _ in range(world)] if rank == 0 else None
        dist.gather(buf, got, dst=0)
    if rank != 0:
        return None
    out = np.full((count, len(RECORD_FIELDS)), -1, np.int64)
    for g in got:
        for row in g.cpu().numpy():
            if row[0] >= 0:
                out[row[0]] = row
    return out


def solve_shard(problems: dict, device: int, configure=None):
    """Solve the instances {id: Problem} of this rank on `device` through the C ABI; returns their records."""
    from . import solver as mcf
    recs = []
    for i, p in problems.items():
        ns = mcf.NetworkSimplex.from_problem(p, device=device)
        if configure is not None:
            configure(ns)
        st = ns.Solve()
        M = ns.GetMetrics()
        if st == mcf.SolverStatus.Optimal:
            recs.append(make_record(i, int(st), M.iterations, ns.GetTotalCost(), ns.flows(), ns.potentials()))
        else:
            recs.append([int(i), int(st), int(M.iterations), 0, 0, 0])
    return recs
