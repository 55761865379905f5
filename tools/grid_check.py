"""Development aid: full team-engine solves of time-expanded grids (no oracle), to find size-dependent problems."""
import os, sys, time, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import mincostflow_b200 as mcf
from mincostflow_b200 import instances
for rows in [int(x) for x in sys.argv[1:]] or [512, 1024]:
    p = instances.grid_time_expanded(rows, rows)
    ns = mcf.NetworkSimplex.from_problem(p); ns.SetOptimizationConfig(mcf.OptimizationConfig())
    ns.set_engine_options(barrier_timeout_s=3.0)
    t = time.time()
    try:
        st = ns.Solve(); M = ns.GetMetrics()
        print(json.dumps(dict(rows=rows, status=int(st), pivots=M.iterations, cost=ns.GetTotalCost() if int(st) == 1 else None, us_per_pivot=round(M.kernel_time_us / max(M.iterations, 1), 3),
                              np=M.pricer_ctas, grid=M.grid_ctas, max_cycle=M.max_cycle, max_stem=M.max_stem, wall=round(time.time() - t, 1))), flush=True)
    except Exception as e:
        print("rows", rows, "ERROR", e, flush=True)
