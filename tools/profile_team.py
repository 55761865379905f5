"""Short team-engine run for ncu source-level captures.  Usage: python tools/profile_team.py [log2n=13] [pivots=3000] [pricing CTAs=auto|0] [engine=team|team_spill]"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import mincostflow_b200 as mcf
from mincostflow_b200 import instances
k = int(sys.argv[1]) if len(sys.argv) > 1 else 13
piv = int(sys.argv[2]) if len(sys.argv) > 2 else 3000
p = instances.netgen8(k)
ns = mcf.NetworkSimplex.from_problem(p)
ns.SetOptimizationConfig(mcf.OptimizationConfig())
ns.set_engine_options(stop_after_pivots=piv, engine=sys.argv[4] if len(sys.argv) > 4 else "team", barrier_timeout_s=5.0,
                      lookahead_blocks=int(sys.argv[3]) if len(sys.argv) > 3 else None)
ns.Solve()
M = ns.GetMetrics()
print(json.dumps(dict(pivots=M.iterations, us_per_pivot=M.kernel_time_us / max(M.iterations, 1), grid=M.grid_ctas, pricers=M.pricer_ctas, wide_flows=M.wide_flows,
                      ph=[round(x / max(M.iterations, 1), 3) for x in M.phase_us])))
