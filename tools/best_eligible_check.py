"""Best Eligible on the BASELINE sizes (config 3: "Block Search vs Best Eligible pivot"): full GPU solves, the CPU oracle timed on
a bounded prefix (a full CPU Best Eligible solve at 2^20 is ~15 h: every pivot scans all 9.4 M arcs), optimality of the GPU
result checked independently, cost cross-checked against the Block Search optimum.
Usage: python tools/best_eligible_check.py 18 20"""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import mincostflow_b200 as mcf
from mincostflow_b200 import instances
from oracle import oracle

large = json.load(open(os.path.join(ROOT, "tests", "golden", "large.json")))
for k in [int(x) for x in sys.argv[1:]]:
    p = instances.netgen8(k)
    ns = mcf.NetworkSimplex.from_problem(p)
    ns.SetPivotRule(mcf.PivotRule.BestEligible).SetOptimizationConfig(mcf.OptimizationConfig())
    t0 = time.perf_counter(); st = ns.Solve(); wall = time.perf_counter() - t0
    M = ns.GetMetrics()
    bad, dual = oracle.validate(p, ns.flows(), ns.potentials(), ns.GetTotalCost())
    cpu_piv = 300 if k >= 18 else 2000
    r, *_ = oracle.solve(p, pivot_rule=oracle.BEST_ELIGIBLE, config=oracle.default_config(), max_pivots=cpu_piv)
    out = dict(instance=p.name, rule="BestEligible", status=int(st), pivots=M.iterations, gpu_solve_s=round(wall, 3), gpu_kernel_s=round(M.kernel_time_us / 1e6, 3),
               gpu_us_per_pivot=round(M.kernel_time_us / M.iterations, 2), pricing_us_per_pivot=round(M.pivot_search_time_us / M.iterations, 2),
               pricing_GBps=round(M.pricing_bytes / M.pivot_search_time_us / 1e3, 1), total_cost=ns.GetTotalCost(),
               equals_block_search_optimum=bool(ns.GetTotalCost() == large[p.name]["total_cost"]), validator_failures=bad, dual_equals_primal=bool(dual == ns.GetTotalCost()),
               cpu_port_us_per_pivot_first_pivots=round(1e6 * r.loop_seconds / r.iterations, 1), cpu_port_sample_pivots=int(r.iterations),
               cpu_port_extrapolated_solve_s=round(r.loop_seconds / r.iterations * M.iterations, 0))
    print(json.dumps(out), flush=True)
