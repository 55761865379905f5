"""mincostflow_b200 - B200-native (sm_100a) network simplex engine behind the solver API of
pdegenhardt/MinCostFlow's `NetworkSimplex` (include/mcfgpu.h).

  solver.py      host-side mirror of the reference's solver surface over the C ABI (ctypes)
  instances.py   NETGEN and time-expanded-grid instance generators, plain-Python DIMACS helpers
  dimacs.py      DimacsReader / SolutionLoader mirror over the native bulk I/O (csrc/mcf_io.cpp)
  csrc/          CUDA kernels (persistent cooperative pivot kernel), C ABI, NETGEN generator
"""
from .solver import (ArgumentException, CompactDigraph, EngineError, GraphBuilder, InvalidOperationException,  # noqa: F401
                     NetworkSimplex, OptimizationConfig, OptimizationFlags, PivotRule, SolverMetrics, SolverStatus,
                     SupplyType, Arc, Node, device_count, load_library, solve_batch, LIB_PATH)
from . import instances  # noqa: F401
from . import dimacs  # noqa: F401

__version__ = "0.1.0"
