"""Writes mid-solve checkpoints of the CPU oracle (oracle/_ref/ckpt_<instance>_<i>.npz, git-ignored, travels with gpurun) so
that bench.py's CPU legs can time windows spread over the whole solve instead of a cheap prefix: the reference's per-pivot
cost grows ~40x over a 2^20 solve (tree depth, cache misses), so a prefix flatters it.
Usage: python tools/make_checkpoints.py 20 [fractions=0.33,0.66]"""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mincostflow_b200 import instances
from oracle import oracle

k = int(sys.argv[1])
fr = [float(x) for x in (sys.argv[2] if len(sys.argv) > 2 else "0.33,0.66").split(",")]
p = instances.netgen8(k)
total = json.load(open(os.path.join(ROOT, "tests", "golden", "large.json")))[p.name]["pivots"]
out_dir = os.path.join(ROOT, "oracle", "_ref")
os.makedirs(out_dir, exist_ok=True)
cfg = oracle.default_config()
st = None
for i, f in enumerate(fr):
    path = os.path.join(out_dir, f"ckpt_{p.name}_{i}.npz")
    if os.path.exists(path) and "--redo" not in sys.argv:          # continue from what is already there
        st = oracle.State.load(path)
        print(f"checkpoint {i}: present (pivot {st.iterations})", flush=True)
        continue
    target = int(total * f)
    nxt = oracle.State(p.n, p.m)
    t = time.time()
    r, *_ = oracle.solve(p, config=cfg, max_pivots=target, resume=st, save=nxt)
    assert r.stopped_early and nxt.iterations == target, (r.iterations, target)
    nxt.save(path)
    print(f"checkpoint {i}: pivot {target} of {total}, {time.time() - t:.1f} s, {os.path.getsize(path) / 1e6:.1f} MB -> {path}", flush=True)
    st = nxt
