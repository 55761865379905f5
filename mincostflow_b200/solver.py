"""Host-side mirror of the reference's solver surface over the C ABI of libmcfgpu.so.

Same names, argument meaning and error behaviour as
  IMinCostFlowSolver        (src/MinCostFlow.Core/IMinCostFlowSolver.cs:8-34)
  NetworkSimplex            (src/MinCostFlow.Core/Lemon/Algorithms/NetworkSimplex.cs:119-210, :416-587)
  GraphBuilder / CompactDigraph (Lemon/Graphs/GraphBuilder.cs:10-104, CompactDigraph.cs:81-127)
so that the parity tests read like the reference's own xUnit tests (src/MinCostFlow.Tests/Lemon/*.cs).
Everything numerical happens in the CUDA engine; this module only marshals flat arrays through ctypes.
There is no CPU fallback: constructing a solver without an sm_100 GPU (or without the built library) raises.
"""
from __future__ import annotations

import ctypes as C
import enum
import os

import numpy as np

_PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_PKG, "libmcfgpu.so")


class ArgumentException(ValueError):
    """System.ArgumentException (NetworkSimplex.cs:155-158)."""


class InvalidOperationException(RuntimeError):
    """System.InvalidOperationException("Solution not optimal") (NetworkSimplex.cs:418-421)."""


class EngineError(RuntimeError):
    """A CUDA / engine failure reported through the C ABI (no reference counterpart)."""

    def __init__(self, code: int, msg: str):
        super().__init__(f"libmcfgpu error {code}: {msg}")
        self.code = code


class SolverStatus(enum.IntEnum):      # Types/SolverStatus.cs:7-34
    NotSolved = 0
    Optimal = 1
    Infeasible = 2
    Unbounded = 3
    Unbalanced = 4


class PivotRule(enum.IntEnum):         # Types/PivotRule.cs:7-40
    FirstEligible = 0
    BestEligible = 1
    BlockSearch = 2
    CandidateList = 3
    AlteringList = 4


class SupplyType(enum.IntEnum):        # Types/SupplyType.cs:7-17
    Geq = 0
    Leq = 1


class OptimizationFlags(enum.IntFlag):  # Algorithms/OptimizationTypes.cs:8-20
    None_ = 0
    AdaptiveBlockSize = 1
    SmallBlocksForDense = 2
    ReducedCostCaching = 4
    CandidateListPivot = 8
    HotColdSplitting = 16
    EarlyTermination = 32


class _CConfig(C.Structure):
    _fields_ = [("flags", C.c_int32), ("max_block_size", C.c_int32), ("min_block_size", C.c_int32),
                ("dense_network_threshold", C.c_int32), ("consecutive_hits_before_adapt", C.c_int32), ("reserved0", C.c_int32),
                ("candidate_list_ratio", C.c_double), ("block_size_growth_factor", C.c_double),
                ("block_size_shrink_factor", C.c_double), ("low_hit_rate_threshold", C.c_double),
                ("high_hit_rate_threshold", C.c_double), ("min_block_size_ratio", C.c_double)]


class _COptions(C.Structure):
    _fields_ = [("supply_type", C.c_int32), ("pivot_rule", C.c_int32), ("auto_configuration", C.c_int32),
                ("optimized_pivot", C.c_int32), ("device", C.c_int32), ("max_ctas", C.c_int32),
                ("lookahead_blocks", C.c_int32), ("engine", C.c_int32), ("simd_width", C.c_int32), ("warm_start", C.c_int32),
                ("stop_after_pivots", C.c_int64),
                ("barrier_timeout_s", C.c_double), ("config", _CConfig)]


class SolverMetrics(C.Structure):      # OptimizationTypes.cs:43-69 + engine counters (mcf_metrics)
    _fields_ = [("iterations", C.c_int64), ("total_arcs_checked", C.c_int64), ("initial_block_size", C.c_int32),
                ("final_block_size", C.c_int32), ("baseline_iterations", C.c_int32), ("pricing_kind", C.c_int32),
                ("average_arcs_checked_per_pivot", C.c_double), ("iteration_ratio", C.c_double),
                ("pivot_search_time_us", C.c_double), ("tree_update_time_us", C.c_double), ("cycle_time_us", C.c_double),
                ("total_solve_time_us", C.c_double), ("kernel_time_us", C.c_double), ("h2d_time_us", C.c_double),
                ("d2h_time_us", C.c_double), ("host_prepass_time_us", C.c_double), ("h2d_bytes", C.c_int64),
                ("d2h_bytes", C.c_int64), ("arcs_priced", C.c_int64), ("pricing_bytes", C.c_int64),
                ("degenerate_pivots", C.c_int64), ("cycle_nodes", C.c_int64), ("moved_nodes", C.c_int64),
                ("max_cycle", C.c_int64), ("max_stem", C.c_int64), ("pricing_rounds", C.c_int64),
                ("config_flags", C.c_int32), ("grid_ctas", C.c_int32), ("degree_cv", C.c_double),
                ("engine", C.c_int32), ("pricer_ctas", C.c_int32), ("stem_exchanges", C.c_int64),
                ("hop_wait_done_us", C.c_double), ("stem_exchange_us", C.c_double), ("ns_per_clock", C.c_double),
                ("phase_us", C.c_double * 16), ("wide_flows", C.c_int32), ("warm_started", C.c_int32)]

    # reference property names
    Iterations = property(lambda s: s.iterations)
    TotalArcsChecked = property(lambda s: s.total_arcs_checked)
    InitialBlockSize = property(lambda s: s.initial_block_size)
    FinalBlockSize = property(lambda s: s.final_block_size)
    AverageArcsCheckedPerPivot = property(lambda s: s.average_arcs_checked_per_pivot)
    BaselineIterations = property(lambda s: s.baseline_iterations)
    IterationRatio = property(lambda s: s.iteration_ratio)
    TotalSolveTimeMicros = property(lambda s: s.total_solve_time_us)
    PivotSearchTimeMicros = property(lambda s: s.pivot_search_time_us)
    TreeUpdateTimeMicros = property(lambda s: s.tree_update_time_us)

    def as_dict(self):
        return {k: (list(getattr(self, k)) if k == "phase_us" else getattr(self, k)) for k, _ in self._fields_}


class OptimizationConfig:
    """OptimizationTypes.cs:25-38 (same defaults)."""

    def __init__(self, Flags=OptimizationFlags.None_, MaxBlockSize=100, MinBlockSize=25, DenseNetworkThreshold=10000,
                 CandidateListRatio=0.1, BlockSizeGrowthFactor=1.2, BlockSizeShrinkFactor=0.8, LowHitRateThreshold=0.05,
                 HighHitRateThreshold=0.3, ConsecutiveHitsBeforeAdapt=3, MinBlockSizeRatio=0.125):
        self.Flags = Flags; self.MaxBlockSize = MaxBlockSize; self.MinBlockSize = MinBlockSize
        self.DenseNetworkThreshold = DenseNetworkThreshold; self.CandidateListRatio = CandidateListRatio
        self.BlockSizeGrowthFactor = BlockSizeGrowthFactor; self.BlockSizeShrinkFactor = BlockSizeShrinkFactor
        self.LowHitRateThreshold = LowHitRateThreshold; self.HighHitRateThreshold = HighHitRateThreshold
        self.ConsecutiveHitsBeforeAdapt = ConsecutiveHitsBeforeAdapt; self.MinBlockSizeRatio = MinBlockSizeRatio

    def _to_c(self) -> _CConfig:
        c = _CConfig()
        c.flags = int(self.Flags); c.max_block_size = self.MaxBlockSize; c.min_block_size = self.MinBlockSize
        c.dense_network_threshold = self.DenseNetworkThreshold; c.consecutive_hits_before_adapt = self.ConsecutiveHitsBeforeAdapt
        c.candidate_list_ratio = self.CandidateListRatio; c.block_size_growth_factor = self.BlockSizeGrowthFactor
        c.block_size_shrink_factor = self.BlockSizeShrinkFactor; c.low_hit_rate_threshold = self.LowHitRateThreshold
        c.high_hit_rate_threshold = self.HighHitRateThreshold; c.min_block_size_ratio = self.MinBlockSizeRatio
        return c


_EXPORTS = ["mcf_api_version", "mcf_device_count", "mcf_default_options", "mcf_create", "mcf_destroy", "mcf_set_arcs",
            "mcf_set_supply", "mcf_set_options", "mcf_solve", "mcf_get_status", "mcf_get_flows", "mcf_get_potentials",
            "mcf_get_flow", "mcf_get_potential", "mcf_get_total_cost", "mcf_get_node_supply", "mcf_get_arc_cost",
            "mcf_get_arc_lower_bound", "mcf_get_arc_upper_bound", "mcf_get_metrics", "mcf_solve_batch", "mcf_solve_batch_concurrent",
            "mcf_pricing_probe", "mcf_validate", "mcf_last_error"]

_lib = None


def load_library():
    """Loads libmcfgpu.so (built by `make -C mincostflow_b200/csrc`); fails loudly when it is missing."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(f"{LIB_PATH} is not built - run `python -c 'import __graft_entry__ as g; g.build()'` "
                              "(the engine has no CPU fallback)")
        lib = C.CDLL(LIB_PATH)
        lib.mcf_last_error.restype = C.c_char_p
        lib.mcf_last_error.argtypes = [C.c_void_p]
        lib.mcf_destroy.restype = None
        lib.mcf_destroy.argtypes = [C.c_void_p]
        lib.mcf_default_options.restype = None
        _lib = lib
    return _lib


def device_count() -> int:
    return int(load_library().mcf_device_count())


class Node(int):
    """Types/Node.cs - an id wrapper."""
    @property
    def Id(self):
        return int(self)


class Arc(int):
    """Types/Arc.cs - an id wrapper."""
    @property
    def Id(self):
        return int(self)


class CompactDigraph:
    """Arc ids in insertion order (CompactDigraph.cs:110-126)."""

    def __init__(self):
        self._n = 0
        self._src = []
        self._tgt = []

    NodeCount = property(lambda s: s._n)
    ArcCount = property(lambda s: len(s._src))

    def AddNode(self):
        self._n += 1
        return Node(self._n - 1)

    def AddArc(self, source, target):
        if not (0 <= int(source) < self._n and 0 <= int(target) < self._n):
            raise ArgumentException("Invalid source or target node")
        self._src.append(int(source)); self._tgt.append(int(target))
        return Arc(len(self._src) - 1)

    def Source(self, arc): return Node(self._src[int(arc)])
    def Target(self, arc): return Node(self._tgt[int(arc)])
    def IsValidArc(self, arc): return 0 <= int(arc) < len(self._src)
    def IsValidNode(self, node): return 0 <= int(node) < self._n

    @staticmethod
    def from_arrays(n, source, target):
        g = CompactDigraph()
        g._n = int(n); g._src = np.ascontiguousarray(source, np.int32); g._tgt = np.ascontiguousarray(target, np.int32)
        return g


class GraphBuilder:
    """GraphBuilder.cs:10-104."""

    def __init__(self):
        self._graph = CompactDigraph()
        self._node_map = {}
        self._next = 0

    def AddNode(self, externalId=None):
        if externalId is None:
            externalId = self._next
            self._next += 1
        if externalId in self._node_map:
            raise ArgumentException(f"Node with ID {externalId} already exists")
        self._node_map[externalId] = self._graph.AddNode()
        return self

    def AddNodes(self, count):
        for _ in range(count):
            self.AddNode()
        return self

    def AddArc(self, sourceId, targetId):
        if sourceId not in self._node_map:
            raise ArgumentException(f"Source node {sourceId} not found")
        if targetId not in self._node_map:
            raise ArgumentException(f"Target node {targetId} not found")
        self._graph.AddArc(self._node_map[sourceId], self._node_map[targetId])
        return self

    def GetNode(self, externalId):
        if externalId not in self._node_map:
            raise ArgumentException(f"Node {externalId} not found")
        return self._node_map[externalId]

    def Build(self):
        return self._graph


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class NetworkSimplex:
    """Drop-in for `MinCostFlow.Core.Lemon.Algorithms.NetworkSimplex` backed by the CUDA engine."""

    def __init__(self, graph, device: int = 0):
        if graph is None:
            raise ArgumentException("graph")                      # ArgumentNullException, NetworkSimplex.cs:121
        self._lib = load_library()
        self._graph = graph
        self._n, self._m = int(graph.NodeCount), int(graph.ArcCount)
        src = np.ascontiguousarray(graph._src, np.int32); tgt = np.ascontiguousarray(graph._tgt, np.int32)
        self._h = C.c_void_p()
        rc = self._lib.mcf_create(C.c_int32(self._n), C.c_int32(self._m), _ptr(src), _ptr(tgt), C.byref(self._h))
        if rc != 0:
            raise EngineError(rc, {-2: "no sm_100 CUDA device (the engine has no CPU fallback)", -1: "invalid graph",
                                   -6: "graph too large"}.get(rc, "mcf_create failed"))
        self._lower = np.zeros(self._m, np.int64)                 # NetworkSimplex.cs:615-617 defaults
        self._upper = np.full(self._m, (2**63 - 1) // 2, np.int64)
        self._cost = np.zeros(self._m, np.int64)
        self._supply = np.zeros(self._n, np.int64)
        self._dirty = True
        self._opt = _COptions()
        self._lib.mcf_default_options(C.byref(self._opt))
        self._opt.device = device
        self._flows = None
        self._pots = None

    def __del__(self):
        try:
            if getattr(self, "_h", None) is not None and self._h.value:
                self._lib.mcf_destroy(self._h)
                self._h = C.c_void_p()
        except Exception:
            pass

    # ------------------------------------------------------------------ errors
    def _check(self, rc):
        if rc == 0:
            return
        msg = (self._lib.mcf_last_error(self._h) or b"").decode()
        if rc == -1:
            raise ArgumentException(msg or "invalid argument")
        if rc == -5:
            raise InvalidOperationException(msg or "Solution not optimal")
        raise EngineError(rc, msg)

    # ------------------------------------------------------------------ setters (fluent, like the reference)
    def SetArcBounds(self, arc, lower, upper):
        if not self._graph.IsValidArc(arc):
            raise ArgumentException("Invalid arc")
        self._lower[int(arc)] = lower; self._upper[int(arc)] = upper; self._dirty = True
        return self

    def SetArcCost(self, arc, cost):
        if not self._graph.IsValidArc(arc):
            raise ArgumentException("Invalid arc")
        self._cost[int(arc)] = cost; self._dirty = True
        return self

    def SetNodeSupply(self, node, supply):
        if not self._graph.IsValidNode(node):
            raise ArgumentException("Invalid node")
        self._supply[int(node)] = supply; self._dirty = True
        return self

    def SetSupplyType(self, type_):
        self._opt.supply_type = int(type_)
        return self

    def SetPivotRule(self, rule):
        self._opt.pivot_rule = int(rule)
        return self

    def EnableOptimizedPivot(self, enable=True, simd_width=None):
        """NetworkSimplex.cs:532.  simd_width = Vector<long>.Count of the host whose pivot sequence the optimized Block Search
        is to reproduce (BlockSearchPivotOptimized.cs:74): 4 on x64 AVX2 (default), 2 on SSE2 / NEON, 0 = not accelerated."""
        self._opt.optimized_pivot = int(bool(enable))
        if simd_width is not None: self._opt.simd_width = int(simd_width)

    def EnableWarmStart(self, enable=True):
        """SURVEY.md 8f-3 (the reference's README.md:17-18 roadmap item; LEMON's re-run semantics network_simplex.h:836-884): a
        Solve() that follows an Optimal Solve() on this object after arc-COST edits only (SetArcCost) starts from that optimal basis
        - tree, arc states, flows kept on the device, potentials recomputed for the new costs - instead of the artificial basis.
        Any other edit (bounds, supplies, supply type) falls back to a cold start.  GetMetrics().warm_started tells which ran."""
        self._opt.warm_start = int(bool(enable))
        return self

    def SetMemoryPool(self, pool):          # stored but never read by the reference (NetworkSimplex.cs:541-544)
        pass

    def EnableOptimizations(self, flags):   # overridden by auto-configuration exactly as in the reference (:549-552 vs :237-241)
        self._opt.config.flags = int(flags)

    def SetOptimizationConfig(self, config: OptimizationConfig):
        if config is None:
            raise ArgumentException("config")
        self._opt.config = config._to_c()
        self._opt.auto_configuration = 0     # NetworkSimplex.cs:560

    def SetAutoConfiguration(self, enable):
        self._opt.auto_configuration = int(bool(enable))

    # bulk setters (what a C# shim would pin and pass in one call)
    def set_arrays(self, lower=None, upper=None, cost=None, supply=None):
        if lower is not None: self._lower[:] = lower
        if upper is not None: self._upper[:] = upper
        if cost is not None: self._cost[:] = cost
        if supply is not None: self._supply[:] = supply
        self._dirty = True
        return self

    def set_engine_options(self, max_ctas=None, lookahead_blocks=None, stop_after_pivots=None, barrier_timeout_s=None, device=None, engine=None):
        if engine is not None: self._opt.engine = {"auto": 0, "flat": 1, "team": 2, "team_spill": 3}.get(engine, engine)
        if max_ctas is not None: self._opt.max_ctas = int(max_ctas)
        if lookahead_blocks is not None: self._opt.lookahead_blocks = int(lookahead_blocks)
        if stop_after_pivots is not None: self._opt.stop_after_pivots = int(stop_after_pivots)
        if barrier_timeout_s is not None: self._opt.barrier_timeout_s = float(barrier_timeout_s)
        if device is not None: self._opt.device = int(device)
        return self

    # ------------------------------------------------------------------ Solve()
    def _push(self):
        if self._dirty:
            self._check(self._lib.mcf_set_arcs(self._h, _ptr(self._lower), _ptr(self._upper), _ptr(self._cost)))
            self._check(self._lib.mcf_set_supply(self._h, _ptr(self._supply)))
            self._dirty = False
        self._check(self._lib.mcf_set_options(self._h, C.byref(self._opt)))

    def Solve(self) -> SolverStatus:
        # CandidateList / AlteringList: the reference throws (NetworkSimplex.cs:884); here they run, defined as LEMON's
        # (network_simplex.h:413-635).  The optimized wrapper still knows only the first three rules (NetworkSimplex.cs:1689-1695).
        if int(self._opt.pivot_rule) > int(PivotRule.BlockSearch) and int(self._opt.optimized_pivot):
            raise NotImplementedError(f"Optimized pivot rule {PivotRule(self._opt.pivot_rule).name} not implemented")
        self._push()
        st = C.c_int32(0)
        self._flows = self._pots = None
        self._check(self._lib.mcf_solve(self._h, C.byref(st)))
        return SolverStatus(st.value)

    @property
    def Status(self) -> SolverStatus:
        st = C.c_int32(0)
        self._check(self._lib.mcf_get_status(self._h, C.byref(st)))
        return SolverStatus(st.value)

    @property
    def SupplyType(self):
        return SupplyType(self._opt.supply_type)

    # ------------------------------------------------------------------ getters
    def _i64(self, fn, idx):
        out = C.c_int64(0)
        self._check(fn(self._h, C.c_int32(int(idx)), C.byref(out)))
        return out.value

    def GetFlow(self, arc): return self._i64(self._lib.mcf_get_flow, arc)
    def GetPotential(self, node): return self._i64(self._lib.mcf_get_potential, node)
    def GetNodeSupply(self, node): return self._i64(self._lib.mcf_get_node_supply, node)
    def GetArcCost(self, arc): return self._i64(self._lib.mcf_get_arc_cost, arc)
    def GetArcLowerBound(self, arc): return self._i64(self._lib.mcf_get_arc_lower_bound, arc)
    def GetArcUpperBound(self, arc): return self._i64(self._lib.mcf_get_arc_upper_bound, arc)

    def GetTotalCost(self):
        out = C.c_int64(0)
        self._check(self._lib.mcf_get_total_cost(self._h, C.byref(out)))
        return out.value

    def GetMetrics(self) -> SolverMetrics:
        m = SolverMetrics()
        self._check(self._lib.mcf_get_metrics(self._h, C.byref(m)))
        return m

    def flows(self) -> np.ndarray:
        if self._flows is None:
            out = np.zeros(self._m, np.int64)
            self._check(self._lib.mcf_get_flows(self._h, _ptr(out)))
            self._flows = out
        return self._flows

    def potentials(self) -> np.ndarray:
        if self._pots is None:
            out = np.zeros(self._n, np.int64)
            self._check(self._lib.mcf_get_potentials(self._h, _ptr(out)))
            self._pots = out
        return self._pots

    def state_after_stop(self):
        """(flow[m], pi[n]) of the basis a bounded solve (stop_after_pivots) stopped at (mcf_get_state_after_stop)."""
        f = np.zeros(self._m, np.int64); p = np.zeros(self._n, np.int64)
        self._check(self._lib.mcf_get_state_after_stop(self._h, _ptr(f), _ptr(p)))
        return f, p

    def device_results(self):
        """(flow pointer, potential pointer, device) of the result arrays in HBM (mcf_get_device_results)."""
        f = C.c_void_p(); p = C.c_void_p(); d = C.c_int32(-1)
        self._check(self._lib.mcf_get_device_results(self._h, C.byref(f), C.byref(p), C.byref(d)))
        return f.value, p.value, d.value

    def Validate(self):
        """SolutionValidator(graph, solver).Validate() on the device (mcf_validate).  Returns (failed-check bits, primal, dual);
        bits == 0 is `IsValid`."""
        bad = C.c_int32(0); primal = C.c_int64(0); dual = C.c_int64(0)
        self._check(self._lib.mcf_validate(self._h, C.byref(bad), C.byref(primal), C.byref(dual)))
        return bad.value, primal.value, dual.value

    def pricing_probe(self, reps=5, flush_l2=True):
        """Stand-alone Best Eligible sweep over all S arcs (mcf_pricing_probe).  Returns (ms per launch, arc, S).
        flush_l2: 0 none, 1 / True overwrite 256 MB then read 256 MB between launches, 2 overwrite only."""
        self._push()
        ms = np.zeros(reps, np.float32)
        arc = C.c_int32(-1); arcs = C.c_int64(0)
        self._check(self._lib.mcf_pricing_probe(self._h, C.c_int32(reps), C.c_int32(int(flush_l2)), _ptr(ms), C.byref(arc), C.byref(arcs)))
        return ms, arc.value, arcs.value

    @classmethod
    def from_problem(cls, p, device: int = 0):
        """`p`: mincostflow_b200.instances.Problem - the per-element setter loop of
        Benchmarks/NetworkSimplexBenchmarks.cs:166-189 as one bulk call."""
        ns = cls(CompactDigraph.from_arrays(p.n, p.source, p.target), device=device)
        return ns.set_arrays(p.lower, p.upper, p.cost, p.supply)


def solve_batch(solvers, devices, per_device: int = 1):
    """mcf_solve_batch[_concurrent]: instance i runs on devices[i % len(devices)], `per_device` solves side by side per GPU."""
    lib = load_library()
    for s in solvers:
        s._push(); s._flows = s._pots = None
    hs = (C.c_void_p * len(solvers))(*[s._h for s in solvers])
    dev = (C.c_int32 * len(devices))(*devices)
    st = (C.c_int32 * len(solvers))()
    rc = lib.mcf_solve_batch_concurrent(hs, C.c_int32(len(solvers)), dev, C.c_int32(len(devices)), C.c_int32(int(per_device)), st)
    if rc != 0:
        raise EngineError(rc, "; ".join((lib.mcf_last_error(s._h) or b"").decode() for s in solvers if lib.mcf_last_error(s._h)))
    return [SolverStatus(x) for x in st]
