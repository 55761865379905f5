"""Debug aid: first pivot at which a bounded GPU solve and the oracle disagree (state after k pivots), per rule.
Usage: python tools/list_rule_debug.py <fixture|netgenK> [maxk]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import mincostflow_b200 as mcf
from mincostflow_b200 import instances
from mincostflow_b200.instances import Problem
from oracle import oracle

name = sys.argv[1]
maxk = int(sys.argv[2]) if len(sys.argv) > 2 else 40
if name.startswith("netgen"):
    p = instances.netgen8(int(name[6:]))
else:
    a = np.load(os.path.join(ROOT, "tests", "golden", "fixtures.npz"))
    p = Problem(int(a[name + ".sup"].shape[0]), int(a[name + ".src"].shape[0]), a[name + ".src"], a[name + ".tgt"], a[name + ".low"], a[name + ".up"],
                a[name + ".cost"], a[name + ".sup"], name)
for rule in (mcf.PivotRule.CandidateList, mcf.PivotRule.AlteringList):
    for k in range(1, maxk + 1):
        ns = mcf.NetworkSimplex.from_problem(p)
        ns.SetPivotRule(rule).SetOptimizationConfig(mcf.OptimizationConfig())
        ns.set_engine_options(stop_after_pivots=k)
        st = ns.Solve()
        M = ns.GetMetrics()
        r, rflow, rpi, tin, _ = oracle.solve(p, pivot_rule=int(rule), auto_config=False, max_pivots=k, trace=k)
        if not r.stopped_early:
            print(rule.name, "oracle finished at", r.iterations); break
        flow, pi = ns.state_after_stop()
        ok = np.array_equal(flow, rflow) and np.array_equal(pi, rpi)
        print(rule.name, "k", k, "ok" if ok else "DIFF", "gpu arcs", M.total_arcs_checked, "oracle arcs", r.total_arcs_checked, "oracle in_arc", int(tin[k - 1]),
              "rounds", M.pricing_rounds, flush=True)
        if not ok:
            d = np.nonzero(flow != rflow)[0][:8]
            print("  flow diff at arcs", d.tolist(), flow[d].tolist(), rflow[d].tolist())
            break
