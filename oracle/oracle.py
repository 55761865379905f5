"""ctypes front-end of the CPU oracle (oracle/ns_oracle.c) and of the vendored-LEMON build (oracle/_ref).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs.  Nothing under mincostflow_b200/ imports this module.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))

FIRST_ELIGIBLE, BEST_ELIGIBLE, BLOCK_SEARCH, CANDIDATE_LIST, ALTERING_LIST = 0, 1, 2, 3, 4
NOT_SOLVED, OPTIMAL, INFEASIBLE, UNBOUNDED, UNBALANCED = 0, 1, 2, 3, 4
GEQ, LEQ = 0, 1
FLAG_ADAPTIVE, FLAG_SMALL_BLOCKS, FLAG_CACHING, FLAG_CANDIDATE, FLAG_HOTCOLD, FLAG_EARLY = 1, 2, 4, 8, 16, 32


class Config(C.Structure):
    _fields_ = [("flags", C.c_int32), ("max_block_size", C.c_int32), ("min_block_size", C.c_int32),
                ("dense_network_threshold", C.c_int32), ("consecutive_hits_before_adapt", C.c_int32), ("_pad", C.c_int32),
                ("candidate_list_ratio", C.c_double), ("block_size_growth_factor", C.c_double),
                ("block_size_shrink_factor", C.c_double), ("low_hit_rate_threshold", C.c_double),
                ("high_hit_rate_threshold", C.c_double), ("min_block_size_ratio", C.c_double)]


class Characteristics(C.Structure):
    _fields_ = [(k, C.c_int32) for k in ("node_count", "arc_count", "max_degree", "source_count", "sink_count",
                                          "transshipment_count", "detected_type", "is_dense", "is_sparse", "is_layered",
                                          "has_uniform_costs", "has_uniform_capacities")] + \
               [(k, C.c_double) for k in ("density", "average_degree", "degree_variance", "degree_cv", "cost_variance",
                                          "average_cost", "cost_cv", "average_capacity", "finite_capacity_ratio")] + \
               [(k, C.c_int64) for k in ("cost_range", "capacity_range", "total_supply", "max_absolute_supply")]


class CState(C.Structure):
    _fields_ = [(k, C.c_void_p) for k in ("parent", "pred", "thread", "rev_thread", "succ_num", "last_succ", "pred_dir", "state", "flow", "pi")] + \
               [("iterations", C.c_int64), ("next_arc", C.c_int32), ("block_size", C.c_int32), ("consecutive_low", C.c_int32),
                ("consecutive_high", C.c_int32)]


class State:
    """Solver state at a pivot boundary (checkpoint / resume of the CPU oracle; see ns_oracle.h)."""
    NODE_I32 = ("parent", "pred", "thread", "rev_thread", "succ_num", "last_succ")

    def __init__(self, n, m):
        self.n, self.m = n, m
        for k in self.NODE_I32:
            setattr(self, k, np.zeros(n + 1, np.int32))
        self.pred_dir = np.zeros(n + 1, np.int8); self.state = np.zeros(m + 2 * n, np.int8)
        self.flow = np.zeros(m + 2 * n, np.int64); self.pi = np.zeros(n + 1, np.int64)
        self.iterations = 0; self.next_arc = 0; self.block_size = 0; self.consecutive_low = 0; self.consecutive_high = 0

    def c(self) -> CState:
        c = CState()
        for k in self.NODE_I32 + ("pred_dir", "state", "flow", "pi"):
            setattr(c, k, getattr(self, k).ctypes.data)
        c.iterations, c.next_arc, c.block_size = self.iterations, self.next_arc, self.block_size
        c.consecutive_low, c.consecutive_high = self.consecutive_low, self.consecutive_high
        return c

    def take(self, c: CState):
        self.iterations, self.next_arc, self.block_size = c.iterations, c.next_arc, c.block_size
        self.consecutive_low, self.consecutive_high = c.consecutive_low, c.consecutive_high

    def save(self, path):
        """Compact file: rev_thread is derived, state / flow are stored sparse (most arcs sit at the lower bound with zero flow)."""
        nz = np.nonzero(self.flow)[0].astype(np.int32)
        ns = np.nonzero(self.state != 1)[0].astype(np.int32)
        np.savez_compressed(path, n=self.n, m=self.m, parent=self.parent, pred=self.pred, thread=self.thread, succ_num=self.succ_num,
                            last_succ=self.last_succ, pred_dir=self.pred_dir, pi=self.pi, flow_idx=nz, flow_val=self.flow[nz],
                            state_idx=ns, state_val=self.state[ns],
                            scalars=np.array([self.iterations, self.next_arc, self.block_size, self.consecutive_low, self.consecutive_high], np.int64))

    @classmethod
    def load(cls, path):
        z = np.load(path)
        st = cls(int(z["n"]), int(z["m"]))
        for k in ("parent", "pred", "thread", "succ_num", "last_succ", "pred_dir", "pi"):
            getattr(st, k)[:] = z[k]
        st.rev_thread[st.thread] = np.arange(st.n + 1, dtype=np.int32)
        st.state[:] = 1; st.state[z["state_idx"]] = z["state_val"]
        st.flow[z["flow_idx"]] = z["flow_val"]
        st.iterations, st.next_arc, st.block_size, st.consecutive_low, st.consecutive_high = (int(x) for x in z["scalars"])
        return st


class Options(C.Structure):
    _fields_ = [("supply_type", C.c_int32), ("pivot_rule", C.c_int32), ("optimized_pivot", C.c_int32),
                ("auto_config", C.c_int32), ("simd_width", C.c_int32), ("collect_phase_times", C.c_int32),
                ("max_pivots", C.c_int64), ("trace_capacity", C.c_int64),
                ("trace_in_arc", C.c_void_p), ("trace_u_out", C.c_void_p), ("config", Config),
                ("resume", C.c_void_p), ("save", C.c_void_p), ("emulate_stackalloc", C.c_int32), ("_pad2", C.c_int32), ("warm", C.c_void_p)]


class Result(C.Structure):
    _fields_ = [(k, C.c_int32) for k in ("status", "pivot_kind", "initial_block_size", "final_block_size", "stopped_early", "_pad")] + \
               [(k, C.c_int64) for k in ("iterations", "total_arcs_checked", "degenerate_pivots", "join_steps", "max_join_steps",
                                         "stem_nodes", "subtree_nodes", "max_subtree_nodes", "total_cost", "art_cost", "sum_supply")] + \
               [(k, C.c_double) for k in ("total_seconds", "loop_seconds", "pricing_seconds", "tree_seconds", "potential_seconds")] + \
               [("config_used", Config), ("characteristics", Characteristics)]


_lib = None


def build(force: bool = False) -> None:
    """Compile oracle/ns_oracle.c (and oracle/_ref when the reference tree is mounted)."""
    so = os.path.join(_HERE, "libns_oracle.so")
    src = os.path.join(_HERE, "ns_oracle.c")
    stale = (not os.path.exists(so)) or os.path.getmtime(so) < max(os.path.getmtime(src), os.path.getmtime(os.path.join(_HERE, "ns_oracle.h")))
    if force or stale:
        subprocess.check_call(["make", "-s", "-C", _HERE, os.path.join(_HERE, "libns_oracle.so")])
    ref_so = os.path.join(_HERE, "_ref", "liblemon_ns.so")
    if os.path.isdir("/root/reference/lemon-1.3.1/lemon") and (force or not os.path.exists(ref_so)
                                                                or os.path.getmtime(ref_so) < os.path.getmtime(os.path.join(_HERE, "lemon_driver.cc"))):
        subprocess.check_call(["make", "-s", "-C", _HERE, "ref"])


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(os.path.join(_HERE, "libns_oracle.so"))
        _lib.ns_oracle_solve.restype = C.c_int
        _lib.ns_oracle_validate.restype = C.c_int
    return _lib


def default_config() -> Config:
    c = Config()
    lib().ns_oracle_default_config(C.byref(c))
    return c


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _arrs(p):
    return (np.ascontiguousarray(p.source, np.int32), np.ascontiguousarray(p.target, np.int32),
            np.ascontiguousarray(p.lower, np.int64), np.ascontiguousarray(p.upper, np.int64),
            np.ascontiguousarray(p.cost, np.int64), np.ascontiguousarray(p.supply, np.int64))


def analyze(p) -> Characteristics:
    src, tgt, lo, up, co, su = _arrs(p)
    ch = Characteristics()
    lib().ns_oracle_analyze(C.c_int(p.n), C.c_int(p.m), _p(src), _p(tgt), _p(lo), _p(up), _p(co), _p(su), C.byref(ch))
    return ch


def select_config(ch: Characteristics) -> Config:
    cfg = Config()
    lib().ns_oracle_select_config(C.byref(ch), C.byref(cfg))
    return cfg


def solve(p, pivot_rule=BLOCK_SEARCH, supply_type=GEQ, auto_config=True, config: Config | None = None,
          optimized_pivot=False, simd_width=4, max_pivots=0, trace=0, phase_times=False, resume: State | None = None,
          save: State | None = None, emulate_stackalloc=False, warm: State | None = None):
    """Returns (Result, flow[int64 m], pi[int64 n], trace_in_arc | None, trace_u_out | None)."""
    src, tgt, lo, up, co, su = _arrs(p)
    o = Options()
    o.supply_type, o.pivot_rule, o.optimized_pivot = supply_type, pivot_rule, int(optimized_pivot)
    o.auto_config = int(auto_config and config is None)
    o.simd_width, o.collect_phase_times, o.max_pivots = simd_width, int(phase_times), int(max_pivots)
    o.config = config if config is not None else default_config()
    tin = tout = None
    if trace:
        tin = np.full(trace, -2, np.int32); tout = np.full(trace, -2, np.int32)
        o.trace_capacity = trace; o.trace_in_arc = tin.ctypes.data; o.trace_u_out = tout.ctypes.data
    o.emulate_stackalloc = int(emulate_stackalloc)
    c_res = c_save = None
    if resume is not None:
        c_res = resume.c(); o.resume = C.addressof(c_res)
    if save is not None:
        c_save = save.c(); o.save = C.addressof(c_save)
    c_warm = None
    if warm is not None:
        c_warm = warm.c(); o.warm = C.addressof(c_warm)
    res = Result()
    flow = np.zeros(p.m, np.int64); pi = np.zeros(p.n, np.int64)
    lib().ns_oracle_solve(C.c_int(p.n), C.c_int(p.m), _p(src), _p(tgt), _p(lo), _p(up), _p(co), _p(su),
                          C.byref(o), C.byref(res), _p(flow), _p(pi))
    if save is not None:
        save.take(c_save)
    return res, flow, pi, tin, tout


def validate(p, flow, pi, total_cost, supply_type=GEQ):
    """SolutionValidator.cs restated; returns (bitmask of failed checks, dual cost)."""
    src, tgt, lo, up, co, su = _arrs(p)
    flow = np.ascontiguousarray(flow, np.int64); pi = np.ascontiguousarray(pi, np.int64)
    dual = C.c_int64(0)
    bad = lib().ns_oracle_validate(C.c_int(p.n), C.c_int(p.m), _p(src), _p(tgt), _p(lo), _p(up), _p(co), _p(su),
                                   C.c_int(supply_type), _p(flow), _p(pi), C.c_int64(int(total_cost)), C.byref(dual))
    return bad, dual.value


# --------------------------------------------------------------------------- vendored LEMON 1.3.1 (oracle/_ref)

_lemon = None


def lemon_available() -> bool:
    return os.path.exists(os.path.join(_HERE, "_ref", "liblemon_ns.so"))


def lemon_solve(p, pivot_rule=BLOCK_SEARCH, supply_type=GEQ):
    """LEMON NetworkSimplex<ListDigraph,int64>::run().  Returns dict(status, cost, seconds, flow, pi).
    Cost/status oracle and CPU baseline only - not a pivot-sequence oracle."""
    global _lemon
    if _lemon is None:
        _lemon = C.CDLL(os.path.join(_HERE, "_ref", "liblemon_ns.so"))
        _lemon.lemon_ns_solve.restype = C.c_int
    src, tgt, lo, up, co, su = _arrs(p)
    flow = np.zeros(p.m, np.int64); pi = np.zeros(p.n, np.int64)
    cost = C.c_int64(0); secs = C.c_double(0)
    st = _lemon.lemon_ns_solve(C.c_int(p.n), C.c_int(p.m), _p(src), _p(tgt), _p(lo), _p(up), _p(co), _p(su),
                               C.c_int(pivot_rule), C.c_int(supply_type), C.byref(cost), C.byref(secs), _p(flow), _p(pi))
    return dict(status=st, cost=cost.value, seconds=secs.value, flow=flow, pi=pi)
