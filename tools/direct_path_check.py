import sys, hashlib, json
sys.path.insert(0, "/root/repo")
import mincostflow_b200 as mcf
from mincostflow_b200 import instances
p = instances.netgen8(19)
res = []
for npr in (1, 0):
    ns = mcf.NetworkSimplex.from_problem(p)
    ns.SetOptimizationConfig(mcf.OptimizationConfig())
    ns.set_engine_options(engine="team", lookahead_blocks=npr, barrier_timeout_s=5.0)
    st = ns.Solve(); M = ns.GetMetrics()
    res.append((int(st), M.iterations, ns.GetTotalCost(), hashlib.sha256(ns.flows().tobytes()).hexdigest(), hashlib.sha256(ns.potentials().tobytes()).hexdigest()))
    print(json.dumps(dict(np=M.pricer_ctas, block=M.initial_block_size, pivots=M.iterations, us=round(M.kernel_time_us / M.iterations, 2), validate=ns.Validate()[0])), flush=True)
print("identical:", res[0] == res[1])
