"""Tiny solves for compute-sanitizer (memcheck / racecheck): team engine on netgen_8_08a and a 24x24 grid (long stems), flat engine
with Best Eligible.  Usage: compute-sanitizer --tool memcheck python tools/sanitize_cmd.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import mincostflow_b200 as mcf
from mincostflow_b200 import instances
from oracle import oracle

for p, rule, eng in ((instances.netgen8(8), mcf.PivotRule.BlockSearch, "team"), (instances.grid_time_expanded(24, 24), mcf.PivotRule.BlockSearch, "team"),
                     (instances.netgen8(8), mcf.PivotRule.BestEligible, None)):
    ns = mcf.NetworkSimplex.from_problem(p)
    ns.SetPivotRule(rule).SetOptimizationConfig(mcf.OptimizationConfig())
    ns.set_engine_options(engine=eng, barrier_timeout_s=60.0, max_ctas=6)
    st = ns.Solve()
    r, fl, pi, _, _ = oracle.solve(p, pivot_rule=int(rule), config=oracle.default_config())
    ok = int(st) == r.status and ns.GetTotalCost() == r.total_cost and np.array_equal(ns.flows(), fl) and np.array_equal(ns.potentials(), pi)
    print(p.name, rule.name, eng, "pivots", ns.GetMetrics().iterations, "match", ok, flush=True)
    assert ok
