#!/usr/bin/env python
"""bench.py -- headline benchmark of the B200 radix hash-join engine.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload B|A|small|zipf0.5|zipf1.0]
                    [--impl ours|reference|reference-cuda]

A "step" is one whole join (histograms + all radix passes of R and S + build/probe + final
reduction) over one synthetic workload.  Metric (BASELINE.json): join throughput in
(|R|+|S|) tuples/s.  N=1 default workload = ETHZ workload B, |R|=|S|=128,000,000 unique
4B-key/4B-payload tuples (the configuration north_star's single-GPU target is quoted on); N>1 =
the same per-GPU shard sizes (weak scaling), radix-sharded with an all-to-all shuffle.

  value         device-resident: inputs already in HBM when the timed region starts
  e2e           through the public host entry (gj_join_aggregate_host): pinned HOST columns in,
                H2D copies + result read-back inside the timed region
  roofline      the dominant kernel (radix scatter pass): algorithmic 16 B/tuple per launch over
                its CUDA-event duration, against MEASURED_PEAKS.json's HBM copy bandwidth
  cpu_baseline  the oracle's multithreaded host radix join (a port: the reference has no CPU
                join) on a bounded sample of the same workload
  reference_cuda  the reference's own CUDA kernels rebuilt for sm_100a (oracle/_ref/bench_ref),
                same inputs via .bin files, run on the same GPU in the same run

--impl reference times the CPU port on the host cores (the reference arm of the driver);
--impl reference-cuda prints the rebuilt reference's line instead.
"""
from __future__ import annotations

import argparse
import json
import os
import re
import statistics
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "join_throughput"
UNIT = "tuples/s"

WORKLOADS = {
    # name: (nR, nS, kind, zipf)
    "B": (128_000_000, 128_000_000, "unique", 0.0),
    "A": (1 << 24, 1 << 28, "fk", 0.0),
    "small": (1 << 20, 1 << 20, "unique", 0.0),
    "zipf0.5": (128_000_000, 128_000_000, "zipf", 0.5),
    "zipf1.0": (128_000_000, 128_000_000, "zipf", 1.0),
    # BASELINE config 5: 2e9 x 2e9 tuples in TOTAL, sharded over the GPUs (strong scaling)
    "cfg5": (2_000_000_000, 2_000_000_000, "unique", 0.0),
}


def workload_name(w):
    nR, nS, kind, z = WORKLOADS[w]
    desc = {"unique": "unique keys both sides", "fk": "ETHZ FK pattern", "zipf": f"Zipf z={z} probe side"}[kind]
    return f"ETHZ workload {w}: |R|={nR}, |S|={nS}, 4B key + 4B payload, {desc}, FK join, count+checksum"


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md: 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks + throttle reasons sampled during the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.index), "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thr = threading.Thread(target=self._read, daemon=True)
            self.thr.start()
        except Exception:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for nm, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


# --------------------------------------------------------------------------------------------
# inputs
# --------------------------------------------------------------------------------------------
def make_host_keys(gj, w, out_R=None, out_S=None):
    """Host key columns of a workload (product generator, multithreaded variants)."""
    nR, nS, kind, z = WORKLOADS[w]
    g = gj.generator
    R = g.create_relation_unique_parallel(nR, nR, 4, out=out_R)
    if kind == "unique":
        S = g.create_relation_unique_parallel(nS, nS, 5, out=out_S)
        expect = nS
    elif kind == "fk":
        S = g.create_relation_unique_parallel(nS, nR, 3, out=out_S)
        expect = nS - (nS - 1) // nR
    else:
        S = g.create_relation_zipf_parallel(nS, nR, z, 7, out=out_S)
        expect = nS - int((S == nR).sum())
    return R, S, expect


def pinned_i32(torch, n):
    t = torch.empty(n, dtype=torch.int32).pin_memory()
    return t, t.numpy()


# --------------------------------------------------------------------------------------------
# arms
# --------------------------------------------------------------------------------------------
def cpu_port_throughput(w, budget_tuples=64_000_000, reps=2):
    """The oracle's multithreaded host radix join on a bounded sample of workload `w`."""
    from oracle import oracle
    nR, nS, kind, z = WORKLOADS[w]
    scale = min(1.0, budget_tuples / (nR + nS))
    mR, mS = max(1024, int(nR * scale)), max(1024, int(nS * scale))
    Rk = oracle.random_unique_gen(mR, mR, 4) if mR <= (1 << 22) else None
    if Rk is None:   # large samples: any permutation will do for a timing sample
        rng = np.random.default_rng(4)
        Rk = rng.permutation(mR).astype(np.int32)
    rng = np.random.default_rng(5)
    if kind == "unique":
        Sk = rng.permutation(mS).astype(np.int32) if mS == mR else rng.integers(0, mR, mS).astype(np.int32)
    elif kind == "fk":
        Sk = rng.integers(0, mR, mS).astype(np.int32)
    else:
        ranks = np.arange(1, mR + 1, dtype=np.float64) ** (-z)
        cdf = np.cumsum(ranks / ranks.sum())
        Sk = np.searchsorted(cdf, rng.random(mS)).clip(0, mR - 1).astype(np.int32)
    ones_r, ones_s = np.ones(mR, np.int32), np.ones(mS, np.int32)
    best = None
    for _ in range(reps):
        res, secs = oracle.join_check(Rk, ones_r, Sk, ones_s, 0, with_time=True)
        best = secs if best is None else min(best, secs)
    return {"value": (mR + mS) / best, "unit": UNIT, "cores": oracle.max_threads(), "kind": "port",
            "sample": f"|R|={mR}, |S|={mS} sample of the same key distribution, best of {reps}, "
                      f"{best * 1e3:.1f} ms (oracle/oracle_join.c orc_join_check, OpenMP)",
            "seconds": best}


def run_reference_cuda(gj, w, steps, timeout_s=420):
    """The reference's CUDA kernels rebuilt for sm_100a, run on identical inputs via .bin files."""
    exe = os.path.join(ROOT, "oracle", "_ref", "bench_ref")
    if not os.path.exists(exe):
        return {"unavailable": "oracle/_ref/bench_ref not built (reference sources absent at build time)"}
    nR, nS, kind, z = WORKLOADS[w]
    if nR >= 128_000_001 or nS >= 128_000_001:
        return {"unavailable": f"the reference dispatches |S|={nS} to its PCIe streaming mode on device 1 "
                               "(hash_join_clustered_probe.cu:2001-2009); only its in-GPU path is in scope"}
    R, S, expect = make_host_keys(gj, w)
    vals, parts, joins, results = [], [], [], None
    with tempfile.TemporaryDirectory(dir="/tmp") as td:
        fr, fs = os.path.join(td, "R.bin"), os.path.join(td, "S.bin")
        R.tofile(fr); S.tofile(fs)
        for _ in range(max(1, steps)):
            try:
                out = subprocess.run([exe, "-b", "7", "-a", "HJC", "-R", str(nR), "-S", str(nS), "--file", "-k", fr, "-l", fs],
                                     capture_output=True, text=True, timeout=timeout_s, cwd=td)
            except subprocess.TimeoutExpired:
                return {"unavailable": f"bench_ref did not finish within {timeout_s}s"}
            txt = out.stdout
            m = re.search(r"(-?\d+) results\s+Without materialization\s+Partition Throughput ([\d.e+]+)\s+"
                          r"Joins Throughput ([\d.e+]+)\s+Total Throughput ([\d.e+]+)", txt)
            if not m:
                return {"unavailable": "bench_ref output not understood: " + (txt[-300:] + out.stderr[-300:]).replace("\n", " | ")}
            results = int(m.group(1))
            parts.append(float(m.group(2)) * 1e6 / 8); joins.append(float(m.group(3)) * 1e6 / 8)
            vals.append(float(m.group(4)) * 1e6 / 8)
    v = statistics.median(vals)
    return {"value": v, "unit": UNIT, "partition_tuples_s": statistics.median(parts), "join_tuples_s": statistics.median(joins),
            "ms_per_step": (nR + nS) / v * 1e3, "results_printed": results, "results_expected": expect,
            "parity": results == (expect if expect < 2**31 else None), "runs": len(vals),
            "what": "reference CUDA kernels (sm_61 launch shapes: 64 CTAs partition, 256 CTAs join) rebuilt for sm_100a, "
                    "'Without materialization' block, data resident, wall-clock around cudaDeviceSynchronize"}


def base_line(args, w, n_gpus):
    return {"metric": METRIC, "unit": UNIT, "n_gpus": n_gpus, "steps": args.steps, "warmup": args.warmup,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int32",
            "data": "synthetic (seeded ETHZ-style generator)",
            "config": {"workload": workload_name(w), "per_gpu_R": WORKLOADS[w][0], "per_gpu_S": WORKLOADS[w][1],
                       "l2": "inputs (>= 2 GB per step) far exceed the 126 MB L2; no flush needed"}}


def reference_arm(args):
    """--impl reference: the CPU port on all host threads (rank 0 only under torchrun)."""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    w = args.workload
    # each step: one bounded sample join
    from oracle import oracle
    oracle.lib()
    vals = []
    cb = None
    for i in range(args.warmup + args.steps):
        cb = cpu_port_throughput(w, budget_tuples=32_000_000, reps=1)
        if i >= args.warmup:
            vals.append(cb["value"])
        if i >= 2 and cb["seconds"] * (args.warmup + args.steps) > 240:   # keep the run within minutes
            break
    v = statistics.median(vals) if vals else cb["value"]
    line = base_line(args, w, args.gpus)
    cb = dict(cb, value=v)
    cb.pop("seconds", None)
    sample_tuples = sum(int(x) for x in re.findall(r"\|[RS]\|=(\d+)", cb["sample"]))
    line.update({"impl": "reference", "value": v, "ms_per_step": sample_tuples / v * 1e3, "cpu_baseline": cb,
                 "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                 "gpu_launches": 0,
                 "note": "the reference has no CPU join (its joinCpu is dead code, hash_join_clustered_probe.cu:2013-2059); "
                         "this arm is the oracle's host radix join (structure of partition-primitives.cu:40-125). "
                         "The reference's CUDA kernels are timed by --impl reference-cuda / key reference_cuda."})
    print(json.dumps(line))


def single_gpu(args):
    import torch
    import __graft_entry__ as ge
    gj = ge.load_package()
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a GPU: the engine has no CPU fallback")
    w = args.workload
    nR, nS, kind, z = WORKLOADS[w]
    torch.cuda.set_device(0)
    pins = [pinned_i32(torch, n) for n in (nR, nR, nS, nS)]
    R, S, expect = make_host_keys(gj, w, pins[0][1], pins[2][1])
    pins[1][1][:] = 1
    pins[3][1][:] = 1          # payload columns of ones, as the reference (hash_join_clustered_probe.cu:1994-1999)
    hRk, hRp, hSk, hSp = (p[0] for p in pins)

    eng = gj.JoinEngine(nR, nS, 0)
    for kv in args.opt:
        k, v = kv.split("=")
        eng.set_option(k, int(v))
    stream = torch.cuda.Stream()
    eng.use_torch_stream(stream)
    dRk, dRp, dSk, dSp = (t.cuda() for t in (hRk, hRp, hSk, hSp))
    torch.cuda.synchronize()

    def check(res):
        if res.matches != expect or res.checksum != expect:
            raise SystemExit(f"WRONG RESULT: matches={res.matches} checksum={res.checksum}, expected {expect}")

    # ---- device-resident throughput ----
    for _ in range(args.warmup):
        check(eng.join_aggregate(dRk, dRp, dSk, dSp))
    sampler = ClockSampler(0)
    sampler.start()
    torch.cuda.synchronize()
    launches0 = gj.kernel_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    tms = []
    e0.record(stream)
    for _ in range(args.steps):
        res = eng.join_aggregate(dRk, dRp, dSk, dSp)
        tms.append(res.timings.as_dict())
    e1.record(stream)
    torch.cuda.synchronize()
    launches = gj.kernel_launch_count() - launches0
    check(res)
    ms_step = e0.elapsed_time(e1) / args.steps
    value = (nR + nS) / (ms_step * 1e-3)

    # ---- roofline of the dominant kernel (radix scatter pass) ----
    peak, peak_src = measured_peak()
    n_of = [min(nR, nS), min(nR, nS), max(nR, nS), max(nR, nS)]    # build p1, build p2, probe p1, probe p2
    per_launch = []
    for t in tms:
        for i, ms in enumerate(t["pass_ms"]):
            if ms > 0:
                per_launch.append((16.0 * n_of[i], ms))
    alg_bytes = statistics.mean(b for b, _ in per_launch)
    avg_ms = statistics.mean(ms for _, ms in per_launch)
    achieved = sum(b for b, _ in per_launch) / sum(ms for _, ms in per_launch) / 1e6   # GB/s
    med = lambda k: statistics.median(t[k] for t in tms)  # noqa: E731
    traffic = None
    tf = os.path.join(ROOT, "profiles", "scatter_dram_bytes.json")
    if os.path.exists(tf):
        try:
            traffic = json.load(open(tf)).get("bytes_per_launch")
        except Exception:
            traffic = None
    roofline = {"bound": "hbm", "kernel": "scatter_kernel (radix scatter pass; %d launches per step)" % sum(1 for x in tms[0]["pass_ms"] if x > 0),
                "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                "peak_source": peak_src, "algorithmic_bytes_per_launch": alg_bytes, "avg_launch_ms": avg_ms,
                "per_phase_ms": {"hist_scan_plan": med("hist_ms"), "scatter_passes": med("part_ms"), "join": med("join_ms"),
                                 "total": med("total_ms")},
                "pipeline_frac": (44.0 if tms[0]["pass2_bits"] else 28.0) * (nR + nS) / (med("total_ms") * 1e-3) / 1e9 / peak}

    # ---- end to end: pinned host columns in, result out ----
    for _ in range(max(1, min(args.warmup, 2))):
        check(eng.join_aggregate_host(hRk, hRp, hSk, hSp))
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    e2e_t = []
    for _ in range(args.steps):
        r2 = eng.join_aggregate_host(hRk, hRp, hSk, hSp)
        e2e_t.append(r2.timings.as_dict())
    torch.cuda.synchronize()
    e2e_ms = (time.perf_counter() - t0) * 1e3 / args.steps
    clocks = sampler.stop()      # sampled across both timed regions (device-resident and end-to-end)
    check(r2)
    e2e = {"value": (nR + nS) / (e2e_ms * 1e-3), "unit": UNIT, "ms_per_step": e2e_ms,
           "h2d_bytes_per_step": 8 * (nR + nS), "d2h_bytes_per_step": 32,
           "h2d_ms": statistics.median(t["h2d_ms"] for t in e2e_t),
           "api": "gj_join_aggregate_host (JoinEngine.join_aggregate_host), pinned host columns"}
    eng.close()
    del dRk, dRp, dSk, dSp
    torch.cuda.empty_cache()

    line = base_line(args, w, 1)
    line.update({"value": value, "ms_per_step": ms_step, "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks,
                 "roofline": roofline,
                 "plan": {"radix_bits": tms[0]["radix_bits"], "pass1_bits": tms[0]["pass1_bits"], "pass2_bits": tms[0]["pass2_bits"]},
                 "checked": f"matches == checksum == {expect} every step"})
    line["config"]["parallelism"] = "1 GPU"
    if not args.no_cpu_baseline:
        cb = cpu_port_throughput(w)
        cb.pop("seconds", None)
        line["cpu_baseline"] = cb
    if not args.no_ref_cuda:
        line["reference_cuda"] = run_reference_cuda(gj, w, 1)
    print(json.dumps(line))


def cfg5_single(args):
    """BASELINE config 5 on ONE GPU (the strong-scaling denominator): 2e9 x 2e9 device-generated
    tuples.  Build partitions are 30.5 K tuples at the 16-bit radix cap, so the join runs its
    multi-chunk steps (8 build chunks x 8 probe chunks per partition) -- reported as measured."""
    import torch
    import __graft_entry__ as ge
    gj = ge.load_package()
    w = args.workload
    n = WORKLOADS[w][0]
    torch.cuda.set_device(0)
    eng = gj.JoinEngine(n, n, 0)
    for kv in args.opt:
        k, v = kv.split("=")
        eng.set_option(k, int(v))
    stream = torch.cuda.Stream()
    eng.use_torch_stream(stream)
    cols = [torch.empty(n, dtype=torch.int32, device="cuda") for _ in range(4)]
    eng.generate_unique(cols[0], cols[1], 0, n, 4, 40)
    eng.generate_unique(cols[2], cols[3], 0, n, 5, 50)
    cols[1].fill_(1); cols[3].fill_(1)
    torch.cuda.synchronize()
    steps, warm = min(args.steps, 3), 1
    for _ in range(warm):
        res = eng.join_aggregate(*cols)
    launches0 = gj.kernel_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    tms = []
    for _ in range(steps):
        res = eng.join_aggregate(*cols)
        tms.append(res.timings.as_dict())
    e1.record(stream)
    torch.cuda.synchronize()
    if res.matches != n or res.checksum != n:
        raise SystemExit(f"WRONG RESULT: {res.matches} {res.checksum}, expected {n}")
    ms_step = e0.elapsed_time(e1) / steps
    peak, peak_src = measured_peak()
    t = tms[-1]
    line = base_line(args, w, 1)
    line.update({"scaling": "strong", "steps": steps, "warmup": warm, "value": 2 * n / (ms_step * 1e-3), "ms_per_step": ms_step,
                 "e2e": None, "gpu_launches": int(gj.kernel_launch_count() - launches0),
                 "roofline": {"bound": "hbm", "kernel": "scatter_kernel", "achieved": 16.0 * 2 * n * 2 / (t["part_ms"] * 1e-3) / 1e9,
                              "peak": peak, "unit": "GB/s", "peak_source": peak_src, "traffic": None,
                              "per_phase_ms": {k: t[k] for k in ("hist_ms", "part_ms", "join_ms", "total_ms")}},
                 "plan": {"radix_bits": t["radix_bits"], "pass1_bits": t["pass1_bits"], "pass2_bits": t["pass2_bits"]},
                 "checked": f"matches == checksum == {n}"})
    line["roofline"]["frac"] = line["roofline"]["achieved"] / peak
    print(json.dumps(line))
    eng.close()


def multi_gpu(args):
    import torch
    import torch.distributed as dist
    import __graft_entry__ as ge
    gj = ge.load_package()
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    w = args.workload
    nR, nS, kind, z = WORKLOADS[w]
    if kind != "unique":
        raise SystemExit("multi-GPU bench supports the unique-key workloads (B, small, cfg5)")
    if args.shuffle == "auto":
        # measured (profiles/README.md, G tuples/s, weak scaling): 2 GPUs pcp 128.8 / pp 124.8 / p2p 111.6;
        # 8 GPUs pcp 406.8 / p2p 352.8 / pp 315.0.  Config 5 has only been measured with p2p and pp at 8 GPUs.
        args.shuffle = "p2p" if (w == "cfg5" and world >= 8) else "pcp"
    strong = (w == "cfg5")
    if strong:                      # fixed total, per-GPU share shrinks with N
        nR, nS = nR // world, nS // world
    NR, NS = nR * world, nS * world
    sj = gj.distributed.ShardedJoin(nR, nS, device=local, mode=args.shuffle, overlap=not args.no_overlap)
    for kv in args.opt:
        k, v = kv.split("=")
        sj.ops.engine.set_option(k, int(v))
    eng = sj.ops.engine
    dev = torch.device("cuda", local)
    cols = [torch.empty(n, dtype=torch.int32, device=dev) for n in (nR, nR, nS, nS)]
    eng.generate_unique(cols[0], cols[1], rank * nR, NR, 4, 40)
    eng.generate_unique(cols[2], cols[3], rank * nS, NS, 5, 50)
    if NR != NS:
        raise SystemExit("unique workloads need |R| == |S|")
    cols[1].fill_(1); cols[3].fill_(1)
    expect = NS
    do_e2e = nR + nS <= 600_000_000
    host = [c.cpu().pin_memory() for c in cols] if do_e2e else None
    torch.cuda.synchronize()

    def step():
        r = sj.join_aggregate(*cols, NR, NS)
        if r.matches != expect or r.checksum != expect:
            raise SystemExit(f"rank {rank}: WRONG RESULT {r.matches} {r.checksum}, expected {expect}")
        return r

    for _ in range(args.warmup):
        step()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    torch.cuda.synchronize(); dist.barrier()
    launches0 = gj.kernel_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        r = step()
    e1.record()
    torch.cuda.synchronize(); dist.barrier()
    launches = gj.kernel_launch_count() - launches0
    ms = torch.tensor([e0.elapsed_time(e1) / args.steps], device=dev)
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms_step = float(ms.item())
    clocks = sampler.stop() if rank == 0 else None

    # end to end: host shards -> device -> sharded join
    def e2e_step():
        for c, h in zip(cols, host):
            c.copy_(h, non_blocking=True)
        torch.cuda.synchronize()
        return step()
    e2 = torch.tensor([float("nan")], device=dev)
    if do_e2e:
        e2e_step()
        torch.cuda.synchronize(); dist.barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            e2e_step()
        torch.cuda.synchronize(); dist.barrier()
        e2 = torch.tensor([(time.perf_counter() - t0) * 1e3 / args.steps], device=dev)
        dist.all_reduce(e2, op=dist.ReduceOp.MAX)
    if rank == 0:
        peak, peak_src = measured_peak()
        line = base_line(args, w, world)
        if strong:
            line["scaling"] = "strong"
            line["config"].update({"per_gpu_R": nR, "per_gpu_S": nS})
        tm = r.phases_ms
        line.update({"value": (NR + NS) / (ms_step * 1e-3), "ms_per_step": ms_step,
                     "e2e": None if not do_e2e else {"value": (NR + NS) / (float(e2.item()) * 1e-3), "unit": UNIT, "ms_per_step": float(e2.item()),
                             "h2d_bytes_per_step": 8 * (NR + NS), "d2h_bytes_per_step": 32 * world,
                             "api": "ShardedJoin.join_aggregate after per-rank H2D of pinned host shards"},
                     "gpu_launches": int(launches) * world, "clocks": clocks,
                     "roofline": {"bound": "hbm", "kernel": "scatter_kernel (local radix passes, rank 0)",
                                  "achieved": 16.0 * (r.local_R + r.local_S) * (2 if tm.get("pass2_bits") else 1) / (tm["part_ms"] * 1e-3) / 1e9 if tm.get("part_ms") else None,
                                  "peak": peak, "unit": "GB/s", "peak_source": peak_src, "traffic": None,
                                  "local_phases_ms": {k: tm.get(k) for k in ("hist_ms", "part_ms", "join_ms", "total_ms")}},
                     "shuffle": {"mode": args.shuffle, "tuples_per_gpu_out": int((nR + nS) * (world - 1) / world),
                                 "scatter_kernel_ms": tm.get("shuffle_scatter_ms"),
                                 "nvlink_out_GBs_per_gpu": (8.0 * (nR + nS) * (world - 1) / world / (tm["shuffle_scatter_ms"] * 1e-3) / 1e9)
                                 if tm.get("shuffle_scatter_ms") else None,
                                 "nvlink_peak_GBs": 770.0, "host_ms_rank0": tm.get("host_ms"), "trace_ms_rank0": tm.get("trace_ms"),
                                 "note": "peer-store scatter: one kernel reads the local shard (8 B/tuple HBM) and stores each run into the destination GPU's HBM over NVLink"},
                     "checked": f"matches == checksum == {expect} every step"})
        if tm.get("pass_ms") and any(tm["pass_ms"][2:]):
            # the receiver's radix passes over S run alone (R's overlap S's shuffle): 16 B/tuple per launch
            ps = [x for x in tm["pass_ms"][2:] if x > 0]
            line["roofline"].update({"kernel": "scatter_kernel (receiver's radix passes over S, rank 0)",
                                     "achieved": 16.0 * r.local_S / (sum(ps) / len(ps) * 1e-3) / 1e9,
                                     "algorithmic_bytes_per_launch": 16.0 * r.local_S, "avg_launch_ms": sum(ps) / len(ps),
                                     "local_phases_ms": {"scatter_R_pass1": tm["pass_ms"][0], "scatter_R_pass2": tm["pass_ms"][1],
                                                         "scatter_S_pass1": tm["pass_ms"][2], "scatter_S_pass2": tm["pass_ms"][3]}})
        if args.shuffle in ("pcp", "pcp2") and tm.get("part_R_ms"):
            line["roofline"].update({"kernel": "pcp source-side radix pass of R (layout + first pass, 16 B/tuple), rank 0",
                                     "achieved": 16.0 * nR / (tm["part_R_ms"] * 1e-3) / 1e9, "algorithmic_bytes_per_launch": 16.0 * nR,
                                     "local_phases_ms": {k: tm.get(k) for k in ("part_R_ms", "copy_R_ms", "recv_R_ms", "part_S_ms", "copy_S_ms", "recv_S_ms", "join_ms")}})
            line["plan"] = {"gpu_bits": world.bit_length() - 1, "local_bits": tm.get("radix_bits"),
                            "source_pass_bits": tm.get("pass1_bits"), "receiver_pass_bits": tm.get("pass2_bits")}
            line["shuffle"]["note"] = ("partition-copy-partition: first radix pass at the source on [gpu | top local bits], whole first-pass "
                                       "partitions bulk-copied (TMA, global->shared->peer global) into the receiver's layout, last pass + join "
                                       "at the receiver; scatter_kernel_ms = the copy kernels of R and S")
        if args.shuffle == "pp" and tm.get("local_R_ms"):
            # dominant HBM-bound kernels of the sharded pipeline: the local phase of one relation =
            # coarse histogram (4 B) + first pass (16 B) + fine counts (8 B) per tuple
            line["roofline"].update({"kernel": "pp local phase of R (hist + first radix pass + fine counts), rank 0",
                                     "achieved": 28.0 * nR / (tm["local_R_ms"] * 1e-3) / 1e9,
                                     "algorithmic_bytes_per_launch": 28.0 * nR,
                                     "local_phases_ms": {k: tm.get(k) for k in ("local_R_ms", "push_R_ms", "local_S_ms", "push_S_ms", "join_ms")}})
            line["plan"] = {"gpu_bits": world.bit_length() - 1, "local_bits": tm.get("radix_bits"),
                            "pass1_bits": tm.get("pass1_bits"), "pass2_bits": tm.get("pass2_bits")}
            line["shuffle"]["note"] = ("partition-then-push: every GPU partitions its own shard on [gpu|local] bits; the last radix pass "
                                       "stores its runs into the destination GPU's final partition buffer over NVLink; "
                                       "scatter_kernel_ms = cursor kernel + pushing pass of R and S")
        if line["shuffle"].get("nvlink_out_GBs_per_gpu"):
            line["shuffle"]["nvlink_frac_of_peer_copy_peak"] = line["shuffle"]["nvlink_out_GBs_per_gpu"] / line["shuffle"]["nvlink_peak_GBs"]
        if line["roofline"]["achieved"]:
            line["roofline"]["frac"] = line["roofline"]["achieved"] / peak
        line["config"].update({"global_R": NR, "global_S": NS, "parallelism": f"radix-sharded over {world} GPUs, {args.shuffle} shuffle"
                               + (", R's push overlapped with S's local pass" if args.shuffle == "pp" else
                                  ", R's copy under S's first pass, S's copy under R's last pass" if args.shuffle == "pcp" else
                                  ", probe side split in two halves, first half joined under the second half's copy" if args.shuffle == "pcp2" else
                                  "" if args.no_overlap or args.shuffle != "p2p" else ", S shuffle overlapped with R's local passes")})
        print(json.dumps(line))
    sj.close()
    dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference", "reference-cuda"])
    ap.add_argument("--workload", default="B", choices=sorted(WORKLOADS))
    ap.add_argument("--shuffle", default="auto", choices=["auto", "p2p", "nccl", "dma", "pp", "pcp", "pcp2"],
                    help="multi-GPU exchange: pp = partition locally, last radix pass pushes into the peers; p2p = peer-store "
                         "shuffle first, local passes at the receiver; pcp = first radix pass at the source, whole first-pass partitions "
                         "bulk-copied over NVLink, last pass at the receiver; auto = pcp (measured best, profiles/README.md)")
    ap.add_argument("--opt", action="append", default=[], help="engine option name=value (repeatable)")
    ap.add_argument("--no-overlap", action="store_true", help="multi-GPU: shuffle and local passes back to back")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-ref-cuda", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        return reference_arm(args)
    if args.impl == "reference-cuda":
        if int(os.environ.get("RANK", "0")) != 0:
            return
        import __graft_entry__ as ge
        gj = ge.load_package()
        line = base_line(args, args.workload, 1)
        ref = run_reference_cuda(gj, args.workload, args.steps)
        line.update({"impl": "reference-cuda", **ref})
        print(json.dumps(line))
        return
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world == 1 and args.gpus == 1 and args.workload == "cfg5":
        return cfg5_single(args)
    if world > 1 or args.gpus > 1:
        if world == 1:
            raise SystemExit("launch multi-GPU runs with torch.distributed.run (see the module docstring)")
        return multi_gpu(args)
    return single_gpu(args)


if __name__ == "__main__":
    main()
