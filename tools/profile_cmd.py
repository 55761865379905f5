"""Short command for ncu captures (B200_PROFILING.md): one bounded Block Search solve, one bounded Best Eligible solve and the
stand-alone pricing sweep on NETGEN-8 instances.  Usage: python tools/profile_cmd.py [log2n=20] [pivots=3000]"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import mincostflow_b200 as mcf
from mincostflow_b200 import instances

k = int(sys.argv[1]) if len(sys.argv) > 1 else 20
piv = int(sys.argv[2]) if len(sys.argv) > 2 else 3000
p = instances.netgen8(k)
out = {}
for rule, cap in ((mcf.PivotRule.BlockSearch, piv), (mcf.PivotRule.BestEligible, max(piv // 10, 50))):
    ns = mcf.NetworkSimplex.from_problem(p)
    ns.SetPivotRule(rule); ns.SetOptimizationConfig(mcf.OptimizationConfig())
    ns.set_engine_options(stop_after_pivots=cap)
    ns.Solve()
    M = ns.GetMetrics()
    out[rule.name] = dict(pivots=M.iterations, kernel_ms=M.kernel_time_us / 1e3, us_per_pivot=M.kernel_time_us / max(M.iterations, 1),
                          price_us=M.pivot_search_time_us / max(M.iterations, 1), cycle_us=M.cycle_time_us / max(M.iterations, 1),
                          update_us=M.tree_update_time_us / max(M.iterations, 1), arcs_priced=M.arcs_priced,
                          pricing_GBps=M.pricing_bytes / max(M.pivot_search_time_us, 1e-9) / 1e3, grid=M.grid_ctas)
ms, arc, S = ns.pricing_probe(reps=6, flush_l2=True)
out["sweep"] = dict(ms=[round(float(x), 4) for x in ms], S=S, GBps=16.0 * S / (float(np.mean(ms[1:])) * 1e-3) / 1e9, arc=arc)
print(json.dumps(out))
