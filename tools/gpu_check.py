"""Development aid: parity + timing sweep on a GPU box (run under gpurun).  Not part of the product."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import mincostflow_b200 as mcf
from mincostflow_b200 import instances
from oracle import oracle

def run(p, rule, cfg=None, auto=False, oracle_too=True, max_ctas=None, stop=None, lookahead=None, engine=None):
    ns = mcf.NetworkSimplex.from_problem(p)
    ns.SetPivotRule(rule)
    if not auto: ns.SetOptimizationConfig(cfg or mcf.OptimizationConfig())
    ns.set_engine_options(max_ctas=max_ctas, stop_after_pivots=stop, lookahead_blocks=lookahead, engine=engine, barrier_timeout_s=3.0)
    t = time.time()
    try:
        st = ns.Solve()
    except Exception as e:
        print("  ENGINE ERROR", e, flush=True); return None
    wall = time.time() - t
    M = ns.GetMetrics()
    out = dict(name=p.name, rule=int(rule), status=int(st), pivots=M.iterations, wall_s=round(wall, 4), kernel_ms=round(M.kernel_time_us / 1e3, 3),
               us_per_pivot=round(M.kernel_time_us / max(M.iterations, 1), 3), price_us=round(M.pivot_search_time_us / max(M.iterations, 1), 3),
               cycle_us=round(M.cycle_time_us / max(M.iterations, 1), 3), update_us=round(M.tree_update_time_us / max(M.iterations, 1), 3),
               grid=M.grid_ctas, np=M.pricer_ctas, wide=M.wide_flows, engine=M.engine, wait_done_us=round(M.hop_wait_done_us / max(M.iterations, 1), 3), stem_x=M.stem_exchanges,
               stem_us=round(M.stem_exchange_us / max(M.iterations, 1), 3), nsclk=round(M.ns_per_clock, 4), ph=[round(x / max(M.iterations, 1), 2) for x in M.phase_us], rounds=M.pricing_rounds, max_cycle=M.max_cycle, max_stem=M.max_stem, kind=M.pricing_kind, flags=M.config_flags)
    if oracle_too:
        oc = None
        if not auto:
            oc = oracle.default_config()
            if cfg is not None:
                oc.flags = int(cfg.Flags)
        r, fl, pi, _, _ = oracle.solve(p, pivot_rule=int(rule), config=oc, auto_config=auto, max_pivots=stop or 0)
        out["oracle_pivots"] = r.iterations; out["oracle_s"] = round(r.loop_seconds, 4)
        if stop:
            out["match"] = bool(M.iterations == r.iterations)
        else:
            out["match"] = bool(int(st) == r.status and M.iterations == r.iterations and (st != 1 or (ns.GetTotalCost() == r.total_cost
                                and np.array_equal(ns.flows(), fl) and np.array_equal(ns.potentials(), pi))))
    print(json.dumps(out), flush=True)
    return out

if __name__ == "__main__":
    what = sys.argv[1] if len(sys.argv) > 1 else "small"
    print("devices", mcf.device_count(), flush=True)
    if what == "small":
        for k in (8, 10, 13):
            p = instances.netgen8(k)
            for rule in (2, 0, 1):
                if rule == 1 and k > 10: continue
                run(p, rule)
            run(p, 2, auto=True)
        p = instances.netgen(13502460, instances.netgen_params(10000, m=30000, sources=100, sinks=100, supply=100000), name="netgen_10k_30k")
        run(p, 2); run(p, 2, auto=True)
        p = instances.netgen8(14); run(p, 2)
        p = instances.netgen8(16); run(p, 2)
        for g in (8, 32, 74, 148):
            run(p, 2, max_ctas=g, oracle_too=False)
    elif what == "team":
        for k in (8, 10, 13, 14, 16):
            p = instances.netgen8(k)
            run(p, 2)
            if k <= 13: run(p, 2, auto=True)
        p = instances.grid_time_expanded(64, 64); run(p, 2)
        p = instances.netgen8(18); run(p, 2, oracle_too=False)
        p = instances.netgen8(20); run(p, 2, oracle_too=False, stop=300000)
    elif what == "quick":
        p = instances.netgen8(13); run(p, 2)
        p = instances.netgen8(16); run(p, 2)
        p = instances.netgen8(20); run(p, 2, oracle_too=False, stop=200000)
    elif what == "big":
        p = instances.netgen8(18); run(p, 2, oracle_too=False)
        p = instances.netgen8(20); run(p, 2, oracle_too=False, stop=200000)
