"""Records what the CPU oracle produces on the 64 instances of BASELINE.json config 5 (NETGEN-8 2^18 nodes, seeds 13502460 + i),
so that the GPU test of the full batch can check every instance bit-exactly (pivot count, cost, sha256 of flow[] and pi[]):
tests/golden/batch18.json.  About 65 s of CPU per instance; `python tools/make_golden_batch18.py [workers=6]`."""
import hashlib, json, os, sys
from concurrent.futures import ProcessPoolExecutor
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
OUT = os.path.join(ROOT, "tests", "golden", "batch18.json")
SEED = 13502460


def one(i):
    from mincostflow_b200 import instances
    from oracle import oracle
    p = instances.netgen8(18, seed=SEED + i)
    r, flow, pi, _, _ = oracle.solve(p, pivot_rule=oracle.BLOCK_SEARCH, config=oracle.default_config())
    return i, dict(seed=SEED + i, status=r.status, pivots=int(r.iterations), total_cost=int(r.total_cost),
                   flow_sha256=hashlib.sha256(flow.tobytes()).hexdigest(), pi_sha256=hashlib.sha256(pi.tobytes()).hexdigest())


if __name__ == "__main__":
    workers = int(sys.argv[1]) if len(sys.argv) > 1 else 6
    data = json.load(open(OUT)) if os.path.exists(OUT) else {}
    todo = [i for i in range(64) if str(i) not in data]
    with ProcessPoolExecutor(workers) as ex:
        for i, e in ex.map(one, todo):
            data[str(i)] = e
            print(i, e["pivots"], e["total_cost"], flush=True)
            json.dump(data, open(OUT, "w"), indent=1, sort_keys=True)
