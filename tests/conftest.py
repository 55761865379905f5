import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with `-m gpu` on the GPU box)")


@pytest.fixture(scope="session")
def golden():
    with open(os.path.join(GOLDEN_DIR, "golden.json")) as f:
        return json.load(f)


@pytest.fixture(scope="session")
def fixture_arrays():
    return np.load(os.path.join(GOLDEN_DIR, "fixtures.npz"))


@pytest.fixture(scope="session")
def load_fixture(fixture_arrays):
    from mincostflow_b200.instances import Problem

    def _load(name):
        a = fixture_arrays
        src = a[name + ".src"]
        return Problem(int(a[name + ".sup"].shape[0]), int(src.shape[0]), src, a[name + ".tgt"], a[name + ".low"],
                       a[name + ".up"], a[name + ".cost"], a[name + ".sup"], name)
    return _load


def lemon_case_problem(golden, case):
    """One of LEMON's 21 min-cost-flow cases (min_cost_flow_test.cc:330-422) as a Problem + expectations."""
    from mincostflow_b200.instances import INF, Problem
    cid, net, low, up, cost, sup, stype, status, total = case
    g = golden["lemon_networks"][net]
    m = len(g["src"])

    def arc_col(key):
        if key == "inf":
            return np.full(m, INF, np.int64)
        if key.startswith("const"):
            return np.full(m, int(key[5:]), np.int64)
        return np.asarray(g["arc_" + key], np.int64)
    p = Problem(g["n"], m, np.asarray(g["src"], np.int32), np.asarray(g["tgt"], np.int32), arc_col(low), arc_col(up),
                arc_col(cost), np.asarray(g["node_" + sup], np.int64), f"lemon_case_{cid}")
    return p, (0 if stype == "GEQ" else 1), {"OPTIMAL": 1, "INFEASIBLE": 2, "UNBOUNDED": 3}[status], total
