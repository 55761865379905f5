"""Host-side instance helpers: DIMACS .min round trip (DimacsReader.cs:60-118 conventions), NETGEN parameters of the reference's
fixture family, the time-expanded grid of BASELINE.json config 4."""
import numpy as np

from mincostflow_b200 import instances


def test_dimacs_round_trip_keeps_arc_ids_and_one_based_nodes():
    p = instances.netgen8(8)
    text = instances.write_dimacs_min(p, ["c test"])
    q = instances.read_dimacs_min(text, "again")
    assert (q.n, q.m) == (p.n, p.m)
    for a in ("source", "target", "lower", "upper", "cost", "supply"):
        assert np.array_equal(getattr(p, a), getattr(q, a)), a
    first_arc = next(l for l in text.splitlines() if l.startswith("a "))
    assert first_arc.split()[1] == str(int(p.source[0]) + 1)            # node ids are 1-based in the file (DimacsReader.cs:88-118)


def test_netgen8_family_parameters():
    """Resources/netgen/netgen_8_08a.min:1-22: m = 8n, sources = sinks = sqrt(n), supply 1000 per source, costs 1..10000, caps 1..1000 (skeleton arcs: up to the supply they must carry)."""
    for k in (8, 10):
        p = instances.netgen8(k)
        n = 1 << k
        assert (p.n, p.m) == (n, 8 * n)
        assert int((p.supply > 0).sum()) == int((p.supply < 0).sum()) == int(round(n ** 0.5))
        assert int(p.supply[p.supply > 0].sum()) == 1000 * int(round(n ** 0.5)) == -int(p.supply[p.supply < 0].sum())
        assert 1 <= p.cost.min() and p.cost.max() <= 10000 and 1 <= p.upper.min() and not p.lower.any()
        assert np.mean(p.upper <= 1000) > 0.9 and p.upper.max() <= p.supply[p.supply > 0].sum()       # skeleton arcs carry a chain's whole supply
    assert not np.array_equal(instances.netgen8(8).cost, instances.netgen8(8, seed=1).cost)


def test_grid_time_expanded_shape():
    p = instances.grid_time_expanded(6, 5, seed=3)
    assert p.n == 30 and p.supply.sum() == 0
    assert (p.target % 5 == p.source % 5 + 1).all()                      # every arc advances one time layer
    assert (np.abs(p.target // 5 - p.source // 5) <= 1).all()            # ... to the same or a neighbouring row
    assert (np.diff(p.source) >= 0).all()                                # arcs grouped by tail node (deterministic ids)
    assert 1 <= p.cost.min() and p.cost.max() <= 10
