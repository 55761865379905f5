"""Full Block Search solves of the two large BASELINE instances on the team engine (tuning aid).  Usage: python tools/full_solve.py [netgen20|grid1024|netgen18] [engine]"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import mincostflow_b200 as mcf
from mincostflow_b200 import instances
which = sys.argv[1] if len(sys.argv) > 1 else "netgen20"
p = instances.grid_time_expanded(1024, 1024) if which == "grid1024" else instances.netgen8(int(which[6:]))
ns = mcf.NetworkSimplex.from_problem(p)
ns.SetOptimizationConfig(mcf.OptimizationConfig())
ns.set_engine_options(engine=sys.argv[2] if len(sys.argv) > 2 else "team")
st = ns.Solve()
M = ns.GetMetrics()
print(json.dumps(dict(inst=which, status=int(st), pivots=M.iterations, us_per_pivot=round(M.kernel_time_us / max(M.iterations, 1), 3), kernel_s=round(M.kernel_time_us / 1e6, 2),
                      wide_flows=M.wide_flows, pricers=M.pricer_ctas, max_stem=M.max_stem, stem_exchanges=M.stem_exchanges, max_cycle=M.max_cycle,
                      moved_per_pivot=round(M.moved_nodes / max(M.iterations, 1), 1))))
