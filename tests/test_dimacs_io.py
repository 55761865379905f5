"""Native bulk DIMACS I/O (csrc/mcf_io.cpp through mincostflow_b200/dimacs.py): DimacsReader.ReadFromStream
(Loaders/DimacsReader.cs:36-147) and SolutionLoader (Loaders/SolutionLoader.cs:60-214) semantics, pinned on a byte-identical
regeneration of the reference's netgen_8_08a.min and on the arrays tools/make_golden.py produced from the reference's fixture
files with the plain-Python reader (conftest.dimacs_dir)."""
import ctypes as C
import glob
import os

import numpy as np
import pytest

from conftest import GOLDEN_DIR
import mincostflow_b200 as mcf
from mincostflow_b200 import dimacs, instances

FIELDS = ("source", "target", "lower", "upper", "cost", "supply")


def _same(p, q):
    assert (p.n, p.m) == (q.n, q.m)
    for a in FIELDS:
        assert np.array_equal(getattr(p, a), getattr(q, a)), a


def test_reference_fixture_files_parse_like_the_reference(load_fixture, golden, dimacs_dir):
    files = sorted(glob.glob(os.path.join(dimacs_dir, "*.min")))
    assert len(files) >= 5
    for path in files:
        name = os.path.splitext(os.path.basename(path))[0]
        p = dimacs.read_from_file(path)
        _same(p, load_fixture(name))                                     # arrays made from the reference's file by the Python reader
        _same(p, instances.read_dimacs_min(path))
        sol = dimacs.load_solution(path[:-4] + ".sol")
        assert sol.OptimalCost == golden["fixtures"][name]["objective"]
        assert sol.ArcFlowsByEndpoints == {(0, 1): 1}                    # `f 1 2 1`: node ids are 1-based in the file


def test_large_text_takes_the_parallel_path_and_keeps_arc_order():
    p = instances.netgen8(14)                                            # 131 072 arcs, ~3 MB of text: split over threads
    text = instances.netgen_dimacs_text(14, 13502460, instances.netgen_params(1 << 14), p)
    assert len(text) > (1 << 20)
    _same(dimacs.read_from_text(text), p)
    _same(dimacs.read_from_text(instances.write_dimacs_min(p, ["c x"] * 3)), p)


def test_grammar_and_errors(dimacs_dir):
    ok = "c comment\n\nx unknown line type is skipped\np min 3 2\nn 1 5\nn 3 -5\nn 1 4\na 1 2 0 10 7\na 2 3 1 +9 -2\r\n"
    p = dimacs.read_from_text(ok)
    assert (p.n, p.m) == (3, 2) and p.supply.tolist() == [4, 0, -5]      # a later `n` line overrides (DimacsReader.cs:92)
    assert p.source.tolist() == [0, 1] and p.target.tolist() == [1, 2] and p.lower.tolist() == [0, 1]
    assert p.upper.tolist() == [10, 9] and p.cost.tolist() == [7, -2]
    empty = dimacs.read_from_text("c nothing\n")
    assert (empty.n, empty.m) == (0, 0)
    for bad, msg in (("p max 3 2\n", "Invalid problem line"), ("p min 3\n", "Invalid problem line"),
                     ("p min 2 1\nn 1\n", "Invalid node line"), ("p min 2 1\na 1 2 0 5\n", "Invalid arc line"),
                     ("p min 2 1\na 1 2 0 5 x\n", "Invalid arc line"), ("p min 2 1\na 1 2 0 99999999999999999999 1\n", "Invalid arc line"),
                     ("p min 2 1\na 1 3 0 5 1\n", "endpoint"), ("p min 2 0\nn 3 1\n", "node id"),
                     ("p min 2 2\na 1 2 0 5 1\n", "declares 2 arcs"), ("a 1 2 0 5 1\n", "no problem line")):
        with pytest.raises(dimacs.FormatException, match=msg):
            dimacs.read_from_text(bad)
    with pytest.raises(OSError):
        dimacs.read_from_file(os.path.join(dimacs_dir, "does_not_exist.min"))


def test_solution_reader_forms(tmp_path):
    f = tmp_path / "a.sol"
    f.write_text("c Gurobi\ns 64\nf 0 10\nf 2 2\nf 3 13\np 0 -3\n")
    s = dimacs.load_solution(str(f))
    assert s.OptimalCost == 64 and s.ArcFlows == {0: 10, 2: 2, 3: 13} and not s.ArcFlowsByEndpoints
    f.write_text("f 1 3 10\nf 2 4 13\n")                                  # flows without an `s` line: "cost not specified" marker
    s = dimacs.load_solution(str(f))
    assert not s.cost_specified and s.ArcFlowsByEndpoints == {(0, 2): 10, (1, 3): 13} and s.ArcFlows[0 * 100000 + 2] == 10


@pytest.mark.gpu
def test_create_from_dimacs_solve_and_write_solution(tmp_path, golden, dimacs_dir):
    """mcf_create_from_dimacs (reader + graph + all setters in one native call) -> Solve -> SaveToFile -> LoadFromFile."""
    lib = mcf.load_library()
    dimacs._lib()
    for name in ("grid_5x5", "netgen_8_08a", "transport_2x3"):
        path = os.path.join(dimacs_dir, name + ".min")
        h = C.c_void_p()
        assert lib.mcf_create_from_dimacs(path.encode(), C.byref(h)) == 0
        st, cost = C.c_int32(), C.c_int64()
        assert lib.mcf_solve(h, C.byref(st)) == 0 and st.value == 1
        assert lib.mcf_get_total_cost(h, C.byref(cost)) == 0 and cost.value == golden["fixtures"][name]["objective"]
        lib.mcf_destroy(h)
        ns = dimacs.solver_from_file(path)
        assert ns.Solve() == mcf.SolverStatus.Optimal
        p = dimacs.read_from_file(path)
        for by_endpoints in (False, True):
            out = str(tmp_path / f"{name}_{int(by_endpoints)}.sol")
            dimacs.save_solution(ns, out, by_endpoints=by_endpoints, with_potentials=not by_endpoints)
            s = dimacs.load_solution(out)
            assert s.OptimalCost == ns.GetTotalCost()
            fl = ns.flows()
            if by_endpoints:
                want = {}
                for e in np.nonzero(fl)[0]:
                    want[(int(p.source[e]), int(p.target[e]))] = int(fl[e])      # parallel arcs: the last line wins, as in the reference's dictionary
                assert s.ArcFlowsByEndpoints == want
            else:
                assert s.ArcFlows == {int(e): int(fl[e]) for e in np.nonzero(fl)[0]}
                lines = open(out).read().splitlines()
                assert lines[0] == f"s {ns.GetTotalCost()}" and sum(l.startswith("p ") for l in lines) == p.n
    unsolved = mcf.NetworkSimplex.from_problem(dimacs.read_from_file(os.path.join(dimacs_dir, "grid_5x5.min")))
    with pytest.raises(mcf.InvalidOperationException):
        dimacs.save_solution(unsolved, str(tmp_path / "x.sol"))
