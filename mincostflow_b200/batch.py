"""Batches of independent instances over several GPUs (BASELINE.json config 5).

A single solve is sequential across pivots and stays on one GPU; a batch shards by instance.  One process per GPU
(`torch.distributed`, NCCL over NVLink; `gloo` in the CPU tests): rank r solves instances r, r + W, r + 2W, ...
and the result records are gathered on rank 0.  No data-path collective exists.

The record of one instance is what SURVEY.md 8e names: {status, pivots, total_cost, flow[m], pi[n]} - 18.9 MB for a
2^18-node NETGEN-8 instance, 1.21 GB for the 64 instances of config 5.  On a GPU the flow / potential arrays are taken
where the solve left them in HBM (`mcf_get_device_results`), packed next to the header on the device, and sent
GPU -> GPU by NCCL; the host never touches them on the sending side.  Rank 0 re-computes the checksums of what arrived and
compares them with the ones the sending rank put into the header."""
from __future__ import annotations

import numpy as np

HEADER_FIELDS = ("instance", "status", "pivots", "total_cost", "flow_checksum", "pi_checksum", "m", "n")
HEADER = len(HEADER_FIELDS)
_GOLD = 0x9E3779B97F4A7C15
_MASK = 0x7FFFFFFFFFFFFFFF


def shard(count: int, world: int, rank: int) -> list[int]:
    """Instance ids owned by `rank` (round robin, like mcf_solve_batch's device assignment)."""
    return list(range(rank, count, world))


def checksum(a: np.ndarray) -> int:
    """Order-sensitive 63-bit checksum of an int64 array (so that permuted results do not collide)."""
    a = np.ascontiguousarray(a, np.int64).view(np.uint64)
    with np.errstate(over="ignore"):
        w = (np.arange(1, a.size + 1, dtype=np.uint64) * np.uint64(_GOLD)) | np.uint64(1)
        return int((a * w).sum(dtype=np.uint64) & np.uint64(_MASK))


def checksum_t(t) -> int:
    """`checksum` of a 1-D int64 torch tensor, computed where the tensor lives (int64 arithmetic wraps like uint64)."""
    import torch
    gold = _GOLD - (1 << 64)                                     # the same bit pattern as a signed 64-bit number
    w = (torch.arange(1, t.numel() + 1, dtype=torch.int64, device=t.device) * gold) | 1
    return int((t * w).sum().item()) & _MASK


class _DeviceArray:
    """Zero-copy view of `count` int64 values at a raw device pointer (`__cuda_array_interface__`)."""

    def __init__(self, ptr: int, count: int):
        self.__cuda_array_interface__ = {"shape": (count,), "typestr": "<i8", "data": (ptr, False), "version": 3, "strides": None}


def device_result_tensors(ns):
    """(flow[m], pi[n]) of a solved NetworkSimplex as torch tensors that alias the engine's own device arrays."""
    import torch
    fptr, pptr, dev = ns.device_results()
    with torch.cuda.device(dev):
        flow = torch.as_tensor(_DeviceArray(fptr, ns._m), device=torch.device("cuda", dev)) if ns._m else torch.empty(0, dtype=torch.int64, device=torch.device("cuda", dev))
        pi = torch.as_tensor(_DeviceArray(pptr, ns._n), device=torch.device("cuda", dev)) if ns._n else torch.empty(0, dtype=torch.int64, device=torch.device("cuda", dev))
    return flow, pi


def pack_record(buf_row, instance: int, status: int, pivots: int, total_cost: int, flow, pi) -> None:
    """Fill one row of the gather buffer: header, flow[m], pi[n] (torch tensors on the row's device)."""
    import torch
    m, n = int(flow.numel()), int(pi.numel())
    hdr = torch.tensor([int(instance), int(status), int(pivots), int(total_cost), checksum_t(flow), checksum_t(pi), m, n], dtype=torch.int64)
    buf_row[:HEADER] = hdr.to(buf_row.device)
    buf_row[HEADER:HEADER + m] = flow
    buf_row[HEADER + m:HEADER + m + n] = pi


def record_width(m_max: int, n_max: int) -> int:
    return HEADER + int(m_max) + int(n_max)


def gather_records(buf, count: int, dist=None):
    """All ranks call with their [per, width] int64 buffer (rows with instance < 0 are padding; `per` = ceil(count / world)
    on every rank).  Rank 0 gets {instance: dict(status, pivots, total_cost, flow, pi, ok)} with the arrays as tensors on
    its device and ok = "the checksums of the arrays that arrived equal the ones in the header"; the others get None."""
    import torch
    world = dist.get_world_size() if dist is not None else 1
    rank = dist.get_rank() if dist is not None else 0
    if dist is None:
        got = [buf]
    else:
        got = [torch.empty_like(buf) for _ in range(world)] if rank == 0 else None
        dist.gather(buf, got, dst=0)
    if rank != 0:
        return None
    out = {}
    for g in got:
        for row in g:
            hdr = row[:HEADER].tolist()
            inst, status, pivots, cost, cf, cp, m, n = hdr
            if inst < 0:
                continue
            flow = row[HEADER:HEADER + m]; pi = row[HEADER + m:HEADER + m + n]
            ok = status != 1 or (checksum_t(flow) == cf and checksum_t(pi) == cp)
            out[int(inst)] = dict(status=int(status), pivots=int(pivots), total_cost=int(cost), flow=flow, pi=pi, ok=bool(ok))
    assert len(out) == count, (sorted(out), count)
    return out


def new_buffer(per: int, width: int, device=None):
    import torch
    buf = torch.zeros((per, width), dtype=torch.int64, device=device)
    buf[:, 0] = -1
    return buf


def solve_shard(problems: dict, device: int, per: int, width: int, configure=None, torch_device=None):
    """Solve the instances {id: Problem} of this rank on `device` through the C ABI and pack their records (GPU path:
    straight from the engine's device arrays).  Returns the [per, width] gather buffer."""
    import torch
    from . import solver as mcf
    tdev = torch_device if torch_device is not None else torch.device("cuda", device)
    buf = new_buffer(per, width, tdev)
    for row, (i, p) in enumerate(problems.items()):
        ns = mcf.NetworkSimplex.from_problem(p, device=device)
        if configure is not None:
            configure(ns)
        st = ns.Solve()
        M = ns.GetMetrics()
        if st == mcf.SolverStatus.Optimal:
            flow, pi = device_result_tensors(ns)
            pack_record(buf[row], i, int(st), M.iterations, ns.GetTotalCost(), flow, pi)
        else:
            buf[row, :4] = torch.tensor([int(i), int(st), int(M.iterations), 0], dtype=torch.int64).to(tdev)
    return buf
