"""Host-side mirror of the reference's problem / solution loaders over the native bulk I/O of libmcfgpu (csrc/mcf_io.cpp):

  DimacsReader.ReadFromFile / ReadFromStream     src/MinCostFlow.Problems/Loaders/DimacsReader.cs:25-147
  SolutionLoader.LoadFromFile / SaveToFile       src/MinCostFlow.Problems/Loaders/SolutionLoader.cs:60-214

The text is parsed in C++ (one parallel pass into the flat arrays the engine uploads); this module only marshals."""
import ctypes as C
from dataclasses import dataclass, field

import numpy as np

from .instances import Problem
from .solver import EngineError, NetworkSimplex, load_library


class FormatException(ValueError):
    """System.FormatException as thrown by DimacsReader (DimacsReader.cs:70, :86, :98)."""


def _lib():
    lib = load_library()
    if not getattr(lib, "_io_ready", False):
        lib.mcf_io_last_error.restype = C.c_char_p
        lib.mcf_dimacs_close.restype = None
        lib.mcf_dimacs_close.argtypes = [C.c_void_p]
        lib.mcf_dimacs_parse.argtypes = [C.c_char_p, C.c_int64, C.POINTER(C.c_void_p)]
        lib.mcf_dimacs_open.argtypes = [C.c_char_p, C.POINTER(C.c_void_p)]
        lib.mcf_dimacs_dims.argtypes = [C.c_void_p, C.POINTER(C.c_int32), C.POINTER(C.c_int32)]
        lib.mcf_dimacs_copy.argtypes = [C.c_void_p] + [C.c_void_p] * 6
        lib.mcf_create_from_dimacs.argtypes = [C.c_char_p, C.POINTER(C.c_void_p)]
        lib.mcf_write_solution.argtypes = [C.c_void_p, C.c_char_p, C.c_int32, C.c_int32]
        lib.mcf_read_solution.argtypes = [C.c_char_p, C.POINTER(C.c_int64), C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p,
                                          C.POINTER(C.c_int32), C.POINTER(C.c_int32)]
        lib._io_ready = True
    return lib


def _raise(lib, rc):
    msg = (lib.mcf_io_last_error() or b"").decode()
    if rc == -10:
        raise FormatException(msg)
    if rc == -11:
        raise OSError(msg)
    raise EngineError(rc, msg)


def _finish(lib, d, name) -> Problem:
    try:
        n, m = C.c_int32(), C.c_int32()
        lib.mcf_dimacs_dims(d, C.byref(n), C.byref(m))
        n, m = n.value, m.value
        src, tgt = np.empty(m, np.int32), np.empty(m, np.int32)
        low, up, cost, sup = np.empty(m, np.int64), np.empty(m, np.int64), np.empty(m, np.int64), np.empty(n, np.int64)
        lib.mcf_dimacs_copy(d, *[a.ctypes.data_as(C.c_void_p) for a in (src, tgt, low, up, cost, sup)])
        return Problem(n, m, src, tgt, low, up, cost, sup, name)
    finally:
        lib.mcf_dimacs_close(d)


def read_from_file(path: str, name: str = "") -> Problem:
    """DimacsReader.ReadFromFile (DimacsReader.cs:25-30)."""
    lib = _lib()
    d = C.c_void_p()
    rc = lib.mcf_dimacs_open(path.encode(), C.byref(d))
    if rc != 0:
        _raise(lib, rc)
    import os
    return _finish(lib, d, name or os.path.basename(path))


def read_from_text(text, name: str = "") -> Problem:
    """DimacsReader.ReadFromStream (DimacsReader.cs:36-147) on text already in memory."""
    lib = _lib()
    raw = text.encode() if isinstance(text, str) else bytes(text)
    d = C.c_void_p()
    rc = lib.mcf_dimacs_parse(raw, len(raw), C.byref(d))
    if rc != 0:
        _raise(lib, rc)
    return _finish(lib, d, name)


def solver_from_file(path: str, device: int = 0) -> NetworkSimplex:
    """ReadFromFile + `new NetworkSimplex(graph)` + the setter loop of NetworkSimplexBenchmarks.cs:166-189 as bulk calls.
    (A C# / C++ host does the same in ONE native call, mcf_create_from_dimacs; this mirror keeps host copies of the arrays
    for its per-element setters, so it goes through them.)"""
    return NetworkSimplex.from_problem(read_from_file(path), device=device)


@dataclass
class Solution:
    """SolutionLoader.Solution (SolutionLoader.cs:18-55), the parts the reference's callers read."""
    OptimalCost: int = 0
    ArcFlows: dict = field(default_factory=dict)                # arc id -> flow (`f ARC_ID FLOW` lines)
    ArcFlowsByEndpoints: dict = field(default_factory=dict)     # (source, target), 0-based -> flow (`f SRC DST FLOW` lines)

    @property
    def cost_specified(self) -> bool:
        return self.OptimalCost != -(2 ** 63)                   # long.MinValue marker, SolutionLoader.cs:165-170


def load_solution(path: str) -> Solution:
    """SolutionLoader.LoadFromFile (SolutionLoader.cs:60-173)."""
    lib = _lib()
    cost, lines, form = C.c_int64(), C.c_int32(), C.c_int32()
    rc = lib.mcf_read_solution(path.encode(), C.byref(cost), 0, None, None, None, C.byref(lines), C.byref(form))
    if rc != 0:
        _raise(lib, rc)
    k = lines.value
    a, b, f = np.empty(k, np.int32), np.empty(k, np.int32), np.empty(k, np.int64)
    rc = lib.mcf_read_solution(path.encode(), C.byref(cost), k, a.ctypes.data_as(C.c_void_p), b.ctypes.data_as(C.c_void_p),
                               f.ctypes.data_as(C.c_void_p), C.byref(lines), C.byref(form))
    if rc != 0:
        _raise(lib, rc)
    s = Solution(OptimalCost=cost.value)
    for x, y, fl in zip(a.tolist(), b.tolist(), f.tolist()):
        if y < 0:
            s.ArcFlows[x] = fl
        else:
            s.ArcFlowsByEndpoints[(x, y)] = fl
            s.ArcFlows[x * 100000 + y] = fl                     # the reference's pseudo arc id (SolutionLoader.cs:131-133)
    return s


def save_solution(ns: NetworkSimplex, path: str, by_endpoints: bool = False, with_potentials: bool = False) -> None:
    """SolutionLoader.SaveToFile (SolutionLoader.cs:186-214) for the solver's current optimal solution."""
    lib = _lib()
    rc = lib.mcf_write_solution(ns._h, path.encode(), C.c_int32(int(by_endpoints)), C.c_int32(int(with_potentials)))
    if rc == -5:
        from .solver import InvalidOperationException
        raise InvalidOperationException("Solution not optimal")
    if rc != 0:
        _raise(lib, rc)
