"""Pins the CPU denominator: ONE full solve of a NETGEN-8 instance by the CPU oracle (the restated
reference, single thread) with its per-bucket cost profile, and ONE full run() of the vendored LEMON 1.3.1
NetworkSimplex (oracle/_ref), both on the host cores of the box this runs on.  Test / measurement infrastructure only.

    python tools/cpu_full_solve.py 20 [oracle|lemon|both] [bucket_pivots]  ->  gpurun_out/r02_cpu_full_<k>_<what>.json

LEMON is timed like the reference's own tools do it (lemon-1.3.1/tools/dimacs-solver.cc:102-129: a timer around run()
only, parsing / graph construction excluded)."""
import json, os, platform, subprocess, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mincostflow_b200 import instances
from oracle import oracle


def host_info():
    info = {"nproc": os.cpu_count(), "machine": platform.machine()}
    try:
        out = subprocess.run(["lscpu"], capture_output=True, text=True, timeout=10).stdout
        for ln in out.splitlines():
            if ln.startswith("Model name"):
                info["cpu_model"] = ln.split(":", 1)[1].strip()
    except Exception:
        pass
    try:
        info["affinity"] = len(os.sched_getaffinity(0))
    except Exception:
        pass
    return info


def main():
    k = int(sys.argv[1]) if len(sys.argv) > 1 else 20
    what = sys.argv[2] if len(sys.argv) > 2 else "both"
    bucket = int(sys.argv[3]) if len(sys.argv) > 3 else 100_000
    p = instances.netgen8(k)
    out = {"instance": p.name, "n": p.n, "m": p.m, "host": host_info(), "when": time.strftime("%Y-%m-%dT%H:%M:%SZ", time.gmtime())}
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    path = os.path.join(ROOT, "gpurun_out", f"r02_cpu_full_{k}_{what}.json")
    if what in ("oracle", "both"):
        cfg = oracle.default_config()
        # the solve cut into buckets (checkpoint / resume is bit-exact and costs a few 10 ms per cut): the total and the cost profile
        prof = []
        st = None
        done = 0
        t0 = time.perf_counter()
        while True:
            nxt = oracle.State(p.n, p.m)
            r2, flow, pi, _, _ = oracle.solve(p, pivot_rule=oracle.BLOCK_SEARCH, config=cfg, max_pivots=done + bucket, resume=st, save=nxt)
            prof.append({"from_pivot": done, "pivots": int(r2.iterations - done), "seconds": r2.loop_seconds})
            done = int(r2.iterations)
            if not r2.stopped_early:
                break
            st = nxt
        wall = time.perf_counter() - t0
        loop = sum(b["seconds"] for b in prof)
        out["oracle_port"] = {"status": r2.status, "pivots": done, "total_cost": int(r2.total_cost), "loop_seconds": loop, "wall_seconds": wall,
                              "pivots_per_s": done / loop, "us_per_pivot": 1e6 * loop / done, "threads": 1,
                              "profile_bucket_pivots": bucket, "profile": prof}
        json.dump(out, open(path, "w"), indent=1)
    if what in ("lemon", "both") and oracle.lemon_available():
        r = oracle.lemon_solve(p, pivot_rule=oracle.BLOCK_SEARCH)
        out["lemon_1_3_1"] = {"status": r["status"], "total_cost": int(r["cost"]), "run_seconds": r["seconds"], "threads": 1,
                              "timed": "NetworkSimplex<ListDigraph,int64>::run(BLOCK_SEARCH) only"}
        json.dump(out, open(path, "w"), indent=1)
    print(json.dumps({k2: v for k2, v in out.items() if k2 != "oracle_port"} | ({"oracle_port": {a: b for a, b in out["oracle_port"].items() if a != "profile"}} if "oracle_port" in out else {})), flush=True)


if __name__ == "__main__":
    main()
