"""Records what the CPU oracle (and the vendored LEMON build) produce on the BASELINE.json NETGEN-8 sizes, so the GPU box
can check full-size solves bit-exactly without spending minutes of CPU time per test:
tests/golden/large.json[name] = {n, m, pivots, total_cost, sha256 of the int64 flow / potential arrays, cpu seconds}.
Block Search, auto-configuration off, OptimizationConfig defaults (the canonical comparator).
Usage: python tools/make_golden_large.py 16 18 20     (2^20 takes ~10 min of CPU)"""
import hashlib
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mincostflow_b200 import instances  # noqa: E402
from oracle import oracle  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden", "large.json")


def main():
    data = json.load(open(OUT)) if os.path.exists(OUT) else {}
    for arg in [a for a in sys.argv[1:] if not a.startswith("--")]:
        if arg.startswith("grid"):
            rows = int(arg[4:])
            p = instances.grid_time_expanded(rows, rows)
        else:
            p = instances.netgen8(int(arg))
        r, flow, pi, _, _ = oracle.solve(p, pivot_rule=oracle.BLOCK_SEARCH, config=oracle.default_config())
        e = dict(n=p.n, m=p.m, status=r.status, pivots=r.iterations, total_cost=r.total_cost, block_size=r.initial_block_size,
                 arcs_checked=r.total_arcs_checked, degenerate=r.degenerate_pivots,
                 flow_sha256=hashlib.sha256(flow.tobytes()).hexdigest(), pi_sha256=hashlib.sha256(pi.tobytes()).hexdigest(),
                 oracle_loop_seconds=round(r.loop_seconds, 3))
        if oracle.lemon_available() and "--no-lemon" not in sys.argv:
            l = oracle.lemon_solve(p)
            e.update(lemon_cost=l["cost"], lemon_seconds=round(l["seconds"], 3))
            assert l["cost"] == r.total_cost
        data[p.name] = e
        print(p.name, e, flush=True)
        with open(OUT, "w") as f:
            json.dump(data, f, indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
