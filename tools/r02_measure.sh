#!/bin/bash
# Round-2 measurement run on one B200 (under gpurun): the bench line of every BASELINE config and the reference arm.
set -u
mkdir -p gpurun_out
python bench.py --steps 2 --warmup 3 --gather > gpurun_out/r02_bench_netgen20.json 2> gpurun_out/r02_bench_netgen20.err; echo "netgen20 rc=$?"
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02_bench_reference_netgen20.json 2> gpurun_out/r02_bench_reference.err; echo "reference rc=$?"
for w in netgen10k netgen16 grid1024 "batch18 --gather"; do
  python bench.py --workload $w --steps 1 --warmup 1 > gpurun_out/r02_bench_${w%% *}.json 2> gpurun_out/r02_bench_${w%% *}.err; echo "$w rc=$?"
done
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r02_bench_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        pk = d.get("pivot_kernel") or {}
        print(f.split("/")[-1], "value", round(d["value"]), d["unit"], "e2e", round(d["e2e"]["value"]), "ms/step", round(d["ms_per_step"], 1), "us/pivot", round(pk.get("us_per_pivot", 0), 3),
              "cpu", round((d.get("cpu_baseline") or {}).get("value", 0)), "roofline", round((d.get("roofline") or {}).get("frac", 0), 3))
    except Exception as e:
        print(f, "unreadable", e)
PY
