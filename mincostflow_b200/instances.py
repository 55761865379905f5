"""Problem instances for the network-simplex path: DIMACS I/O and the NETGEN / grid generators.

Mirrors what the reference feeds its solver with:
  * `read_dimacs_min` / `read_dimacs_sol` follow `MinCostFlow.Problems/Loaders/DimacsReader.cs:60-147`
    (1-based ids -> 0-based, arcs keep file order = arc id) and `Loaders/SolutionLoader.cs:60-173`.
  * `netgen` drives `csrc/netgen.c` (the reference ships NETGEN instances but no generator;
    `Resources/netgen/netgen_8_08a.min:1-22` gives the parameters of the "NETGEN-8" family).
  * `grid_time_expanded` is BASELINE.json's config 4 (`SURVEY.md` section 8d item 4).
"""
from __future__ import annotations

import ctypes
import os
import subprocess
from dataclasses import dataclass

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
INF = (2**63 - 1) // 2          # NetworkSimplex.cs:127


@dataclass
class Problem:
    """Flat arrays in arc-id / node-id order (what `NetworkSimplex`'s per-element setters receive)."""
    n: int
    m: int
    source: np.ndarray      # int32[m]
    target: np.ndarray      # int32[m]
    lower: np.ndarray       # int64[m]
    upper: np.ndarray       # int64[m]
    cost: np.ndarray        # int64[m]
    supply: np.ndarray      # int64[n]
    name: str = ""


def read_dimacs_min(path_or_text: str, name: str = "") -> Problem:
    if "\n" in path_or_text:
        text = path_or_text
    else:
        with open(path_or_text, "r") as f:
            text = f.read()
        name = name or os.path.basename(path_or_text)
    n = m = 0
    sup = {}
    arcs = []
    for line in text.splitlines():
        tok = line.split()
        if not tok:
            continue
        if tok[0] == "p":
            if len(tok) != 4 or tok[1] != "min":
                raise ValueError(f"Invalid problem line: {line}")
            n, m = int(tok[2]), int(tok[3])
        elif tok[0] == "n":
            if len(tok) != 3:
                raise ValueError(f"Invalid node line: {line}")
            sup[int(tok[1]) - 1] = int(tok[2])
        elif tok[0] == "a":
            if len(tok) != 6:
                raise ValueError(f"Invalid arc line: {line}")
            arcs.append((int(tok[1]) - 1, int(tok[2]) - 1, int(tok[3]), int(tok[4]), int(tok[5])))
    a = np.array(arcs, dtype=np.int64).reshape(-1, 5)
    supply = np.zeros(n, np.int64)
    for k, v in sup.items():
        supply[k] = v
    return Problem(n, len(arcs), a[:, 0].astype(np.int32), a[:, 1].astype(np.int32), a[:, 2].copy(),
                   a[:, 3].copy(), a[:, 4].copy(), supply, name)


def read_dimacs_sol(path: str):
    """Returns (objective, {(tail, head): flow}) with 0-based node ids; `s` and `f` lines only."""
    obj = None
    flows = []
    with open(path) as f:
        for line in f:
            tok = line.split()
            if not tok:
                continue
            if tok[0] == "s":
                obj = int(float(tok[1]))
            elif tok[0] == "f":
                flows.append((int(tok[1]) - 1, int(tok[2]) - 1, int(tok[3])))
    return obj, flows


def write_dimacs_min(p: Problem, header_lines=()) -> str:
    out = list(header_lines)
    out.append(f"p min {p.n} {p.m}")
    for i in np.nonzero(p.supply)[0]:
        out.append(f"n {i + 1} {p.supply[i]}")
    for e in range(p.m):
        out.append(f"a {p.source[e] + 1} {p.target[e] + 1} {p.lower[e]} {p.upper[e]} {p.cost[e]}")
    return "\n".join(out) + "\n"


# --------------------------------------------------------------------------- NETGEN

_gen_lib = None


def _libgen():
    global _gen_lib
    if _gen_lib is None:
        so = os.path.join(_HERE, "csrc", "libmcfgen.so")
        if not os.path.exists(so):
            subprocess.check_call(["gcc", "-O2", "-shared", "-fPIC", "-o", so, os.path.join(_HERE, "csrc", "netgen.c")])
        _gen_lib = ctypes.CDLL(so)
        _gen_lib.mcfgen_netgen.restype = ctypes.c_int
    return _gen_lib


def netgen_params(n: int, m: int | None = None, sources: int | None = None, sinks: int | None = None,
                  supply: int | None = None, mincost=1, maxcost=10000, mincap=1, maxcap=1000):
    """Parameters of the NETGEN-8 family exactly as the reference's fixtures are parameterised
    (netgen_8_08a.min:1-22): m = 8n, sources = sinks = round(sqrt n), supply = 1000 * sources, no
    transshipment sources/sinks, 100 % skeleton arcs at max cost, 100 % capacitated."""
    s = int(round(n ** 0.5)) if sources is None else sources
    t = s if sinks is None else sinks
    return [n, s, t, 8 * n if m is None else m, mincost, maxcost, 1000 * s if supply is None else supply,
            0, 0, 100, 100, mincap, maxcap]


def netgen(seed: int, parms, name: str = "") -> Problem:
    lib = _libgen()
    n, m = int(parms[0]), int(parms[3])
    P = (ctypes.c_int64 * 13)(*[int(x) for x in parms])
    tail = np.zeros(m + 16, np.int32); head = np.zeros(m + 16, np.int32)
    cost = np.zeros(m + 16, np.int64); cap = np.zeros(m + 16, np.int64); sup = np.zeros(n, np.int64)
    na = ctypes.c_int64(0)
    ptr = lambda a: a.ctypes.data_as(ctypes.c_void_p)
    rc = lib.mcfgen_netgen(ctypes.c_int64(seed), P, ctypes.c_int64(m + 16), ctypes.byref(na),
                           ptr(tail), ptr(head), ptr(cost), ptr(cap), ptr(sup))
    if rc != 0:
        raise ValueError(f"netgen failed with code {rc}")
    k = na.value
    return Problem(n, k, tail[:k] - 1, head[:k] - 1, np.zeros(k, np.int64), cap[:k].copy(), cost[:k].copy(), sup,
                   name or f"netgen_n{n}_m{m}_seed{seed}")


def netgen8(log2n: int, seed: int = 13502460) -> Problem:
    n = 1 << log2n
    return netgen(seed, netgen_params(n), name=f"netgen_8_{log2n:02d}a" + ("" if seed == 13502460 else f"_s{seed}"))


def netgen_dimacs_text(problem_no: int, seed: int, parms, p: Problem) -> str:
    """The generator's own banner, so that regenerated fixtures can be compared byte for byte."""
    names = ["Number of nodes:      ", "Source nodes:         ", "Sink nodes:           ", "Number of arcs:       ",
             "Minimum arc cost:     ", "Maximum arc cost:     ", "Total supply:         "]
    L = ["c NETGEN flow network generator (C version)", "c  Problem %2d input parameters" % problem_no,
         "c  ---------------------------", "c   Random seed:          %10d" % seed]
    L += ["c   %s%10d" % (nm, v) for nm, v in zip(names, parms[:7])]
    L += ["c   Transshipment -", "c     Sources:            %10d" % parms[7], "c     Sinks:              %10d" % parms[8],
          "c   Skeleton arcs -", "c     With max cost:      %10d%%" % parms[9], "c     Capacitated:        %10d%%" % parms[10],
          "c   Minimum arc capacity: %10d" % parms[11], "c   Maximum arc capacity: %10d" % parms[12], "c",
          "c  *** Minimum cost flow ***", "c"]
    return write_dimacs_min(p, L)


# --------------------------------------------------------------------------- time-expanded grid (config 4)

def _splitmix64(state: np.ndarray) -> np.ndarray:
    z = (state + np.uint64(0x9E3779B97F4A7C15))
    z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
    z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
    return z ^ (z >> np.uint64(31))


def grid_time_expanded(rows: int, cols: int, seed: int = 42, supply_per_row: int = 10) -> Problem:
    """rows x cols grid, node id = r*cols + c (ProblemGenerator.cs:121); column = time layer.
    Arcs (r,c)->(r,c+1) [hold-over] and (r,c)->(r+-1,c+1); cost U[1,10] (ProblemGenerator.cs:134,143);
    capacity = per-layer supply; sources in column 0, sinks in column cols-1, balanced.
    RNG = counter-based splitmix64(seed, arc index): the reference's System.Random(42) stream is not
    reproducible without .NET."""
    r = np.arange(rows, dtype=np.int64)[:, None]
    c = np.arange(cols - 1, dtype=np.int64)[None, :]
    src_l, tgt_l = [], []
    for dr in (0, -1, 1):
        rr = r + dr
        ok = (rr >= 0) & (rr < rows) & (c >= 0)
        s_ = (r * cols + c) + 0 * rr
        t_ = rr * cols + c + 1
        src_l.append(s_[ok]); tgt_l.append(t_[ok])
    src = np.concatenate(src_l); tgt = np.concatenate(tgt_l)
    order = np.lexsort((tgt, src))                      # arcs grouped by tail node, deterministic ids
    src, tgt = src[order], tgt[order]
    m = src.shape[0]
    with np.errstate(over="ignore"):
        h = _splitmix64(np.uint64(seed) * np.uint64(0x2545F4914F6CDD1D) + np.arange(m, dtype=np.uint64))
    cost = (h % np.uint64(10)).astype(np.int64) + 1
    cap_total = rows * supply_per_row
    upper = np.full(m, cap_total, np.int64)
    supply = np.zeros(rows * cols, np.int64)
    supply[np.arange(rows) * cols] = supply_per_row
    supply[np.arange(rows) * cols + cols - 1] = -supply_per_row
    return Problem(rows * cols, m, src.astype(np.int32), tgt.astype(np.int32), np.zeros(m, np.int64), upper, cost,
                   supply, f"grid_te_{rows}x{cols}_s{seed}")
