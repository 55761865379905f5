#!/usr/bin/env python
"""bench.py - BASELINE.json's metric on its own config: one network-simplex solve of NETGEN-8 2^20 nodes / 2^23 arcs
(Block Search, auto-configuration off = the canonical comparator of SURVEY.md A.3) per step, per GPU.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
                    [--workload netgen20|netgen18|netgen16|netgen10k|grid1024|batch18]      (BASELINE.json configs 3, -, 2, 1, 4, 5)

One JSON line on stdout (rank 0).  Keys follow the driver contract:
  value         pivots/s, whole job, device-timed: CUDA events around the persistent pivot kernel on the stream it is launched
                on, inputs resident in HBM (max over ranks of the summed kernel time)
  e2e           the same metric through the reference-facing API (NetworkSimplex.Solve() over the C ABI of libmcfgpu.so) with
                HOST buffers: host pre-pass, H2D of the instance, kernel, D2H of flows + potentials are inside the timed region
  roofline      the stand-alone Best Eligible pricing sweep (the HBM-bound kernel of the path), timed live with CUDA events
  cpu_baseline  the CPU oracle (C restatement of the reference's NetworkSimplex.cs) on a bounded sample of the same workload
  pivot_kernel  in-kernel phase split of the persistent kernel (pricing / cycle / update) and its own pricing GB/s
N > 1 (torchrun, one rank per GPU): every rank solves its own NETGEN instance of the same size (seed + rank) - a single
solve is sequential across pivots and does not shard - and the full result records {status, pivots, cost,
flow[m], pi[n]} are gathered on rank 0 over NCCL straight from the engine's device arrays (mincostflow_b200/batch.py).
Warm-up steps are BOUNDED solves of the same instance (the first 200 000 pivots; `config.warmup_kind`): they warm the
context, the allocations and the instruction caches, which is all a warm-up does for a solver whose timed step is one
20-second kernel; timed steps are full solves.
`--impl reference` times the CPU oracle (oracle/, the restated reference) on the box's host cores instead.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

SEED = 13502460
WORKLOADS = {
    # name: (log2 n or generator tag, instances per rank per step, CPU sample: pivots of the prefix window [0 = the full solve])
    "netgen20": (20, 1, 600_000),         # BASELINE.json config 3 - the headline
    "netgen18": (18, 1, 300_000),
    "netgen16": (16, 1, 0),               # config 2
    "netgen10k": ("10k", 1, 0),           # config 1 (NETGEN 10 000 nodes / 30 000 arcs)
    "grid1024": ("grid", 1, -400_000),    # config 4 (1024 x 1024 time-expanded grid); negative: a plain prefix (no checkpoints recorded)
    "batch18": (18, 8, 300_000),          # config 5: 64 instances of 2^18 nodes = 8 per GPU at 8 GPUs
}
MID_WINDOW = {20: 40_000, 18: 100_000}    # pivots timed from each mid-solve checkpoint (oracle/_ref/ckpt_*.npz)
WARMUP_PIVOTS = 200_000                   # warm-up steps stop after this many pivots


class MissingCheckpoints(RuntimeError):
    pass


def host_info():
    info = {"nproc": os.cpu_count()}
    try:
        for ln in subprocess.run(["lscpu"], capture_output=True, text=True, timeout=10).stdout.splitlines():
            if ln.startswith("Model name"):
                info["cpu_model"] = ln.split(":", 1)[1].strip()
    except Exception:
        pass
    return info


def cpu_sample(p, k, prefix, scale=1.0):
    """Times the CPU oracle on a bounded sample of ONE solve of `p`: the first `prefix` pivots and, when the checkpoints
    written by tools/make_checkpoints.py travelled with the repo, a window from each of them (the reference's per-pivot cost
    grows by more than an order of magnitude over a solve; a prefix alone flatters it).  The mean is weighted by how much of
    the solve each window stands for.  Returns (pivots/s, description, per-window list)."""
    import glob
    from oracle import oracle
    cfg = oracle.default_config()
    wins = []
    prefix_only = prefix < 0
    prefix = int(abs(prefix) * scale)
    r, *_ = oracle.solve(p, pivot_rule=oracle.BLOCK_SEARCH, config=cfg, max_pivots=prefix)
    wins.append({"from_pivot": 0, "pivots": int(r.iterations), "seconds": r.loop_seconds})
    if prefix == 0:
        return r.iterations / r.loop_seconds, "the full solve", wins
    if prefix_only:
        return r.iterations / r.loop_seconds, f"the first {prefix} pivots of one solve only (flatters the CPU: its per-pivot cost grows over a solve)", wins
    ck = sorted(glob.glob(os.path.join(ROOT, "oracle", "_ref", f"ckpt_{p.name}_*.npz")))
    rec = (recorded_cpu() or {}).get(p.name)
    if not ck or not rec:
        # a prefix alone flatters the CPU by 2x (its per-pivot cost grows over the solve): no silent fallback
        raise MissingCheckpoints(f"oracle/_ref/ckpt_{p.name}_*.npz not found: run `python -c 'import __graft_entry__ as g; g.build()'` "
                                 f"(or tools/make_checkpoints.py {k} 0.33,0.66,0.85) once - about 15 CPU-minutes - or pass --no-cpu")
    total = rec["pivots"]
    mid = max(int(MID_WINDOW[k] * scale), 2000)
    for path in ck:
        st = oracle.State.load(path)
        start = st.iterations
        r, *_ = oracle.solve(p, pivot_rule=oracle.BLOCK_SEARCH, config=cfg, max_pivots=start + mid, resume=st)
        wins.append({"from_pivot": int(start), "pivots": int(r.iterations - start), "seconds": r.loop_seconds})
    if scale >= 1.0:
        # the reference itself zero-fills a stackalloc int[n] on every stem re-hang (NetworkSimplex.cs:1085); the port hoists that
        # scratch (conservative).  One mid-solve window is repeated with the zero-fill, for the record only.
        st = oracle.State.load(ck[len(ck) // 2]); start = st.iterations
        r, *_ = oracle.solve(p, pivot_rule=oracle.BLOCK_SEARCH, config=cfg, max_pivots=start + mid // 4, resume=st, emulate_stackalloc=True)
        wins[1 + len(ck) // 2]["us_per_pivot_with_reference_stackalloc"] = 1e6 * r.loop_seconds / max(r.iterations - start, 1)
    # each window stands for the stretch of the solve up to the next window's start ...
    starts = [w["from_pivot"] for w in wins] + [total]
    est_seconds = sum(w["seconds"] / w["pivots"] * (starts[i + 1] - starts[i]) for i, w in enumerate(wins))
    # ... or, better, when the cost profile of one full solve on this pool's hosts is on record (profiles/r02_cpu_full_20.json): every
    # 100 000-pivot bucket of that profile is scaled by (cost per pivot measured now) / (cost per pivot on record) of the window it
    # belongs to.  A step function through four windows misses how steeply the last sixth of the solve climbs (115 -> 210 us per pivot).
    prof = recorded_profile(p.name)
    if prof:
        def rec_cost(a, b):                                            # recorded seconds per pivot over pivots [a, b)
            sec = piv = 0.0
            for bk in prof:
                lo, hi = max(a, bk["from_pivot"]), min(b, bk["from_pivot"] + bk["pivots"])
                if hi > lo:
                    sec += bk["seconds"] * (hi - lo) / bk["pivots"]; piv += hi - lo
            return sec / piv if piv else None
        ratios = []
        for w in wins:
            rc = rec_cost(w["from_pivot"], w["from_pivot"] + w["pivots"])
            ratios.append((w["seconds"] / w["pivots"]) / rc if rc else 1.0)
            w["cost_vs_recorded_profile"] = ratios[-1]
        est2 = 0.0
        for bk in prof:
            i = max(j for j in range(len(wins)) if starts[j] <= bk["from_pivot"])
            est2 += bk["seconds"] * ratios[i]
        for w in wins:
            w["step_estimate_seconds"] = est_seconds
        est_seconds = est2
    how = "each scaling its part of the recorded cost profile of one full solve (profiles/r02_cpu_full_20.json)" if prof else "each weighted by the stretch of the solve it stands for"
    desc = (f"{len(wins)} windows of one solve ({', '.join(str(w['pivots']) + ' pivots from pivot ' + str(w['from_pivot']) for w in wins)}; "
            f"mid-solve windows resume checkpoints the same code wrote), {how}; {total} pivots")
    return total / est_seconds, desc, wins


def workload_name(w):
    k, per, _ = WORKLOADS[w]
    if k == "10k":
        return f"NETGEN 10000 nodes / 30000 arcs (100 sources, 100 sinks, supply 100000), seed {SEED}+i, Block Search, auto-configuration off"
    if k == "grid":
        return "1024 x 1024 time-expanded grid (splitmix64 seed 42+i), Block Search, auto-configuration off"
    return (f"NETGEN-8 2^{k} nodes / 2^{k + 3} arcs, seed {SEED}+i, Block Search, auto-configuration off"
            + (f", {per} instances per GPU per step" if per > 1 else ""))


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms during the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx = gpu_index
        self.proc = None
        self.lines = []
        self.skip = 0

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.idx}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._pump, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def mark(self):
        """The timed region starts here: earlier samples are dropped."""
        self.skip = len(self.lines)

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for ln in (self.lines[self.skip:] or self.lines[-1:]):
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def make_instances(workload, rank):
    from mincostflow_b200 import instances
    k, per, _ = WORKLOADS[workload]
    if k == "10k":
        return [instances.netgen(SEED + rank, instances.netgen_params(10000, m=30000, sources=100, sinks=100, supply=100000), name="netgen_10k_30k")]
    if k == "grid":
        return [instances.grid_time_expanded(1024, 1024, seed=42 + rank)]
    return [instances.netgen8(k, seed=SEED + rank * per + i) for i in range(per)]


def measured_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json, burst copy)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def recorded_full_cpu(name):
    """One FULL solve of the instance by the oracle port and by LEMON 1.3.1, recorded on a B200 box's host (tools/cpu_full_solve.py,
    profiles/r02_cpu_full_20.json): the yardstick the sampled estimate is checked against."""
    path = os.path.join(ROOT, "profiles", "r02_cpu_full_20.json")
    if not os.path.exists(path):
        return {}
    with open(path) as f:
        d = json.load(f)
    if d.get("instance") != name:
        return {}
    o = d.get("oracle_port", {}); l = d.get("lemon_1_3_1", {})
    return {"recorded_full_solve": {"oracle_port_s": o.get("loop_seconds"), "oracle_port_pivots_per_s": o.get("pivots_per_s"),
                                    "lemon_1_3_1_run_s": l.get("run_seconds"), "pivots": o.get("pivots"), "host": d.get("host"),
                                    "where": "host cores of a B200 box of this pool, profiles/r02_cpu_full_20.json"}}


def recorded_profile(name):
    path = os.path.join(ROOT, "profiles", "r02_cpu_full_20.json")
    if not os.path.exists(path):
        return None
    with open(path) as f:
        d = json.load(f)
    return d.get("oracle_port", {}).get("profile") if d.get("instance") == name else None


def recorded_cpu():
    path = os.path.join(ROOT, "tests", "golden", "large.json")
    if not os.path.exists(path):
        return None
    with open(path) as f:
        return json.load(f)


# ----------------------------------------------------------------------------------------------- reference arm

def run_reference(args):
    """The reference's own CPU path (restated: oracle/ns_oracle.c) on a bounded sample: the first P pivots of the workload."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    k, per, sample = WORKLOADS[args.workload]
    p = make_instances(args.workload, 0)[0]
    vals, secs = [], []
    sample_txt = ""
    # every timed step is the same bounded sample of one solve, sized so that the whole run ends within a few minutes:
    # a quarter of the cpu_baseline sample per step at the driver's 20 steps (warm-up steps: a twentieth; a CPU needs none)
    scale = min(1.0, 5.0 / max(args.steps, 1))
    for it in range(args.warmup + args.steps):
        t0 = time.perf_counter()
        v, sample_txt, wins = cpu_sample(p, k, sample, scale=scale if it >= args.warmup else 0.05)
        if it >= args.warmup:
            vals.append(v); secs.append(time.perf_counter() - t0)
    value = float(np.mean(vals))
    total_t = sum(secs); times = secs
    sample_txt += f"; one {workload_name(args.workload)} instance per step"
    out = {"impl": "reference", "metric": "pivots_per_s", "value": value, "unit": "pivots/s", "n_gpus": args.gpus, "steps": args.steps,
           "warmup": args.warmup, "ms_per_step": 1e3 * total_t / max(len(times), 1), "higher_is_better": True, "scaling": "weak",
           "vs_baseline": None, "dtype": "int64", "data": "synthetic",
           "config": {"workload": workload_name(args.workload)},        # (the sample is described in cpu_baseline.sample: same `config` as our arm)
           "cpu_baseline": {"value": value, "unit": "pivots/s", "cores": 1, "kind": "port", "sample": sample_txt, **host_info(),
                            **recorded_full_cpu(p.name)},
           "e2e": {"value": value, "unit": "pivots/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(out), flush=True)


# ----------------------------------------------------------------------------------------------- our arm

def run_ours(args):
    import torch
    import mincostflow_b200 as mcf
    from mincostflow_b200 import batch

    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    dist = None
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    if mcf.device_count() <= 0:
        raise SystemExit("bench.py: no sm_100 GPU visible - the engine has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)

    k, per, sample = WORKLOADS[args.workload]
    probs = make_instances(args.workload, rank)
    solvers = []
    for p in probs:
        ns = mcf.NetworkSimplex.from_problem(p, device=local_rank)
        ns.SetPivotRule(mcf.PivotRule.BlockSearch)
        ns.SetOptimizationConfig(mcf.OptimizationConfig())
        solvers.append(ns)

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    width = batch.record_width(max(p.m for p in probs), max(p.n for p in probs))

    def step(bounded=False):
        """One pass of the hot path over this rank's batch; returns (pivots, kernel_us, launches, h2d, d2h).  `bounded`: warm-up."""
        piv = kus = h2d = d2h = 0
        launches = 0
        for ns in solvers:
            ns.set_engine_options(stop_after_pivots=WARMUP_PIVOTS if bounded else 0)
            ns._dirty = True                                   # host buffers are marshalled and uploaded again every step
        if per > 1:                                            # a batch: `args.concurrency` solves side by side on this GPU
            t0 = time.perf_counter()
            sts = mcf.solve_batch(solvers, [local_rank], per_device=args.concurrency)
            batch_us = (time.perf_counter() - t0) * 1e6
        for i, ns in enumerate(solvers):
            st = sts[i] if per > 1 else ns.Solve()
            M = ns.GetMetrics()
            assert bounded or st == mcf.SolverStatus.Optimal, st
            piv += M.iterations; kus += M.kernel_time_us if per == 1 else 0; h2d += M.h2d_bytes; d2h += M.d2h_bytes; launches += 1
        if per > 1:
            kus = batch_us                                     # solves overlap: the figure is the host-timed span of the batch
        return piv, kus, launches, h2d, d2h

    def gather():
        """NCCL gather of the full result records {status, pivots, cost, flow[m], pi[n]} on rank 0, packed on the
        device from the arrays the solves left in HBM.  Returns (milliseconds on this rank, records verified on rank 0)."""
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        buf = batch.new_buffer(per, width, dev)
        for i, ns in enumerate(solvers):
            flow, pi = batch.device_result_tensors(ns)
            batch.pack_record(buf[i], rank * per + i, int(ns.Status), ns.GetMetrics().iterations, ns.GetTotalCost(), flow, pi)
        got = batch.gather_records(buf, world * per, dist=dist)
        torch.cuda.synchronize()
        ms = (time.perf_counter() - t0) * 1e3
        if rank == 0:
            assert all(v["ok"] and v["status"] == 1 for v in got.values()), {i: (v["status"], v["ok"]) for i, v in got.items()}
        return ms, (len(got) if got is not None else 0)

    # nvidia-smi is started before the warm-up (its start-up takes 0.2-0.3 s and holds driver locks, which would land inside a
    # sub-second timed region); only the samples taken after the timed region begins are used
    sampler = ClockSampler(local_rank); sampler.start()
    for _ in range(args.warmup):
        step(bounded=True)
    if dist is not None:
        # warm-up of the collective itself: NCCL opens its send / receive channels on the first gather (0.5 s at N = 4, measured);
        # the bounded warm-up solves above end NotSolved and have no record to send, so a one-row dummy goes instead
        wb = torch.zeros((1, 16), dtype=torch.int64, device=dev)
        wl = [torch.empty_like(wb) for _ in range(world)] if rank == 0 else None
        dist.gather(wb, wl, dst=0)
    barrier()
    sampler.mark()
    t0 = time.perf_counter()
    tot_piv = tot_kus = tot_h2d = tot_d2h = tot_launch = 0
    gather_ms = 0.0; gathered_n = 0
    for _ in range(args.steps):
        piv, kus, launches, h2d, d2h = step()
        tot_piv += piv; tot_kus += kus; tot_h2d += h2d; tot_d2h += d2h; tot_launch += launches
        if dist is not None or args.gather:
            g_ms, gathered_n = gather()
            gather_ms += g_ms
    barrier()
    wall = time.perf_counter() - t0
    clocks = sampler.stop()

    # max over ranks of the timed region, sums of the work
    stats = torch.tensor([wall, tot_kus, float(tot_piv), float(tot_h2d), float(tot_d2h), float(tot_launch)], dtype=torch.float64, device=dev)
    if dist is not None:
        mx = stats.clone(); dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        sm = stats.clone(); dist.all_reduce(sm, op=dist.ReduceOp.SUM)
        wall, kus_max = mx[0].item(), mx[1].item()
        piv_all, h2d_all, d2h_all, launch_all = sm[2].item(), sm[3].item(), sm[4].item(), sm[5].item()
    else:
        kus_max = tot_kus; piv_all, h2d_all, d2h_all, launch_all = tot_piv, tot_h2d, tot_d2h, tot_launch

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return

    M = solvers[0].GetMetrics()
    peak, peak_src = measured_peak()
    # roofline kernel: stand-alone Best Eligible pricing sweep over all S arcs of the same instance, L2 flushed between launches
    # (a 256 MB buffer is overwritten, then a second one is read so that the first one's dirty lines are not written back inside the
    # timed launch; the figure after an overwrite-only flush is reported next to it)
    ms, arc, S = solvers[0].pricing_probe(reps=12, flush_l2=1)
    sweep_ms = float(np.mean(ms[2:]))
    achieved = 16.0 * S / (sweep_ms * 1e-3) / 1e9
    ms_w, _, _ = solvers[0].pricing_probe(reps=12, flush_l2=2)
    achieved_w = 16.0 * S / (float(np.mean(ms_w[2:])) * 1e-3) / 1e9
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "roofline_traffic.json")
    if os.path.exists(tpath):
        with open(tpath) as f:
            traffic = json.load(f).get(args.workload)
    value = piv_all / (kus_max * 1e-6)
    e2e = piv_all / wall
    out = {
        "metric": "pivots_per_s", "value": value, "unit": "pivots/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * wall / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int64",
        "data": "synthetic",
        "config": {"workload": workload_name(args.workload), "instances_per_gpu_per_step": per, "pivots_per_step_rank0": tot_piv // args.steps,
                   "warmup_kind": f"bounded solves of the same instance (first {WARMUP_PIVOTS} pivots); timed steps are full solves",
                   "value_timing": "CUDA events around the pivot kernel" if per == 1 else "host wall clock around mcf_solve_batch_concurrent (the solves overlap on the device)",
                   "l2": "inputs larger than L2 (arc arrays 16 B x %d arcs; every step re-uploads them)" % S,
                   "grid_ctas": M.grid_ctas, "pricing_kind": M.pricing_kind, "block_size": M.initial_block_size},
        "solve_ms": {"kernel": kus_max / 1e3 / args.steps / per, "end_to_end": 1e3 * wall / args.steps / per},
        "e2e": {"value": e2e, "unit": "pivots/s", "h2d_bytes_per_step": int(h2d_all / args.steps), "d2h_bytes_per_step": int(d2h_all / args.steps)},
        "gpu_launches": int(launch_all),
        "result_gather": ({"ms_per_step_rank0": gather_ms / args.steps, "records_on_rank0": gathered_n, "bytes_per_record": 8 * width,
                           "what": "NCCL gather of {status, pivots, cost, flow[m], pi[n]} from every rank's device arrays; checksums re-verified on rank 0"}
                          if (dist is not None or args.gather) else None),
        "clocks": clocks,
        "roofline": {"bound": "hbm", "kernel": "ns_price_sweep_kernel (Best Eligible full scan, 16 B/arc x S arcs per launch)",
                     "achieved": achieved, "peak": peak, "peak_source": peak_src, "unit": "GB/s", "frac": achieved / peak,
                     "l2_flush": "256 MB overwritten + 256 MB read between launches (L2 = 126 MB)",
                     "achieved_after_overwrite_only_flush": achieved_w,
                     "traffic": traffic, "ms_per_launch": sweep_ms, "arcs_per_launch": int(S)},
        "pivot_kernel": {"us_per_pivot": M.kernel_time_us / max(M.iterations, 1),
                         "pricing_us_per_pivot": M.pivot_search_time_us / max(M.iterations, 1),
                         "cycle_us_per_pivot": M.cycle_time_us / max(M.iterations, 1),
                         "update_us_per_pivot": M.tree_update_time_us / max(M.iterations, 1),
                         "arcs_priced_per_pivot": M.arcs_priced / max(M.iterations, 1),
                         # in-kernel probes of the team engine (clock64 deltas of one pricing CTA and of the first owner CTA), us per pivot
                         "pricer_cta_us": {nm: M.phase_us[i] / max(M.iterations, 1) for nm, i in (
                             ("price_and_post_ENTER", 1), ("collect_ENTER", 2), ("stage_next_block_arcs", 3), ("wait_DONE", 6), ("finish_staging", 0),
                             ("gather_node_records_post_GATHERED", 7), ("wait_and_gather_CYC", 4), ("decide", 5))} if M.engine == 2 else None,
                         "owner_cta_us": {nm: M.phase_us[i] / max(M.iterations, 1) for nm, i in (
                             ("wait_and_collect_ENTER", 9), ("scan_slice", 10), ("reduce_and_post_CYC", 11), ("gather_CYC_and_decide", 12),
                             ("cycle_node_update", 15), ("relabel", 13), ("post_DONE", 14))} if M.engine == 2 else None,
                         "pricing_GBps_in_kernel": M.pricing_bytes / max(M.pivot_search_time_us, 1e-9) / 1e3},
    }
    # CPU baseline (oracle port) on a bounded sample + the GPU on the very same sample
    if not args.no_cpu and world == 1:
        t0 = time.perf_counter()
        cpu_v, sample_txt, wins = cpu_sample(probs[0], k, sample)
        cpu_s = time.perf_counter() - t0
        out["cpu_baseline"] = {"value": cpu_v, "unit": "pivots/s", "cores": 1, "kind": "port", "sample": sample_txt + " (same instance)",
                               "seconds": cpu_s, "windows": wins, **host_info(),
                               "prefix_only_value": wins[0]["pivots"] / wins[0]["seconds"], **recorded_full_cpu(probs[0].name)}
        full = out["cpu_baseline"].get("recorded_full_solve")
        if full and full.get("oracle_port_pivots_per_s"):
            out["cpu_baseline"]["sampling_error_vs_recorded_full_solve"] = cpu_v / full["oracle_port_pivots_per_s"] - 1.0
        if sample:
            ns = solvers[0]
            ns.set_engine_options(stop_after_pivots=abs(sample))
            ns._dirty = True
            t0 = time.perf_counter(); ns.Solve(); gw = time.perf_counter() - t0
            Ms = ns.GetMetrics()
            ns.set_engine_options(stop_after_pivots=0)
            ns._dirty = True
            out["cpu_baseline"]["gpu_same_sample"] = {"pivots": Ms.iterations, "kernel_pivots_per_s": Ms.iterations / (Ms.kernel_time_us * 1e-6),
                                                      "e2e_pivots_per_s": Ms.iterations / gw}
    print(json.dumps(out), flush=True)
    if dist is not None:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="netgen20", choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--concurrency", type=int, default=4, help="batch workloads: solves side by side per GPU")
    ap.add_argument("--gather", action="store_true", help="run the result-record packing / gather also at N = 1 (it always runs at N > 1)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
