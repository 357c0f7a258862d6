#!/usr/bin/env python
"""bench.py -- headline benchmark of the B200 radix hash-join engine.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload B|A|small|zipf0.5|zipf1.0]
                    [--impl ours|reference|reference-cuda]

A "step" is one whole join (histograms + all radix passes of R and S + build/probe + final
reduction) over one synthetic workload.  Metric (BASELINE.json): join throughput in
(|R|+|S|) tuples/s.  N=1 default workload = ETHZ workload B, |R|=|S|=128,000,000 unique
4B-key/4B-payload tuples (the configuration north_star's single-GPU target is quoted on); N>1 =
the same per-GPU shard sizes (weak scaling), radix-sharded with an all-to-all shuffle.

  value         device-resident: inputs already in HBM when the timed region starts; every step is
                checked: match count and 64-bit payload checksum against the closed form of the
                generated key sets (payload = mix(key), computed here with torch, independent of
                the engine and of the oracle)
  e2e           through the public host entry (gj_join_aggregate_host): pinned HOST columns in,
                H2D copies + result read-back inside the timed region
  roofline      the dominant kernel (radix scatter pass): algorithmic 16 B/tuple per launch over
                its CUDA-event duration, against MEASURED_PEAKS.json's HBM copy bandwidth
  cpu_baseline  the oracle's multithreaded host radix join (a port: the reference has no CPU
                join) on a bounded sample of the same workload
  materialize   the materialising join (exact-size pair output) on the same inputs
  config5       BASELINE config 5 (2e9 x 2e9 tuples in total, strong scaling): the per-GPU share
                2e9/N, device-generated, 5 timed steps after 2 warm-ups, with the speed-up over the 1-GPU run
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


PAY_SEED_R, PAY_SEED_S = 40, 50      # payload = mix32(key ^ f(seed)), bit for bit gj_payload_of_key (csrc/kernels.cuh)
_M32 = 0xFFFFFFFF


def _mix32_t(x):
    """mix32 of csrc/kernels.cuh on an int64 torch tensor holding uint32 values."""
    x = x ^ (x >> 16)
    x = (x * 0x7feb352d) & _M32
    x = x ^ (x >> 15)
    x = (x * 0x846ca68b) & _M32
    return x ^ (x >> 16)


def payload_i64(torch, keys, seed, lo=0, hi=None):
    """payload_of_key(key, seed) of keys[lo:hi] as SIGNED values in an int64 tensor."""
    k = keys[lo:hi].to(torch.int64) & _M32
    c = (seed * 0xC2B2AE35 + 0x27D4EB2F) & _M32
    v = _mix32_t(k ^ c)
    return torch.where(v >= (1 << 31), v - (1 << 32), v)


def fill_payloads(torch, keys, pays, seed, chunk=1 << 26):
    for lo in range(0, keys.numel(), chunk):
        pays[lo:lo + chunk] = payload_i64(torch, keys, seed, lo, lo + chunk).to(torch.int32)


def closed_form(torch, probe_keys, n_build_keys, chunk=1 << 26):
    """Expected (matches, checksum) when the build side holds every key of [0, n_build_keys) exactly
    once, from the PROBE key column alone: a probe tuple matches iff 0 <= key < n_build_keys and then
    contributes payload(key, 40) * payload(key, 50) (int64 arithmetic wraps = mod 2^64)."""
    m, c = 0, 0
    for lo in range(0, probe_keys.numel(), chunk):
        k = probe_keys[lo:lo + chunk]
        hit = (k >= 0) & (k.to(torch.int64) < n_build_keys)
        prod = payload_i64(torch, k, PAY_SEED_R) * payload_i64(torch, k, PAY_SEED_S)
        m += int(hit.sum().item())
        c = (c + int(torch.where(hit, prod, torch.zeros_like(prod)).sum().item())) & 0xFFFFFFFFFFFFFFFF
    return m, c


def pinned_i32(torch, n):
    t = torch.empty(n, dtype=torch.int32).pin_memory()
    return t, t.numpy()


# --------------------------------------------------------------------------------------------
# arms
# --------------------------------------------------------------------------------------------
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


def config_of(w, n_gpus):
    """The `config` object of a line: a function of (workload, N) only, so that the GPU arm and the
    reference arm of one driver run print the same one."""
    nR, nS, kind, z = WORKLOADS[w]
    strong = (w == "cfg5")
    per_R, per_S = (nR // n_gpus, nS // n_gpus) if strong else (nR, nS)
    return {"workload": workload_name(w), "per_gpu_R": per_R, "per_gpu_S": per_S,
            "global_R": per_R * n_gpus, "global_S": per_S * n_gpus,
            "parallelism": "1 GPU" if n_gpus == 1 else f"radix-sharded over {n_gpus} GPUs (top radix bits = GPU id), "
                           + ("strong scaling" if strong else "weak scaling: the same shard sizes on every GPU"),
            "l2": ("inputs (>= 2 GB per step) far exceed the 126 MB L2; no flush needed" if 8 * (per_R + per_S) > (1 << 29) else
                   "inputs fit the 126 MB L2 (a latency-bound case by design, not the headline); no flush")}


def base_line(args, w, n_gpus):
    return {"metric": METRIC, "unit": UNIT, "n_gpus": n_gpus, "steps": args.steps, "warmup": args.warmup,
            "higher_is_better": True, "scaling": "strong" if w == "cfg5" else "weak", "vs_baseline": None, "dtype": "int32",
            "data": "synthetic (seeded ETHZ-style generator; payload = mix32(key))",
            "config": config_of(w, n_gpus)}


def host_threads():
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


CPU_SAMPLE_MAX = 128_000_000     # tuples per side the CPU arm joins per step (workload B in full)


def cpu_join_sample(w, n_gpus=1):
    """Inputs of the CPU arm: the stated workload in full where one side has at most 128 M tuples
    (workload B: all of it), else a sample with the same |R|:|S| ratio and key distribution."""
    cfg = config_of(w, n_gpus)
    nR, nS = cfg["global_R"], cfg["global_S"]
    kind, z = WORKLOADS[w][2], WORKLOADS[w][3]
    scale = min(1.0, CPU_SAMPLE_MAX / max(nR, nS))
    mR, mS = max(1024, int(nR * scale)), max(1024, int(nS * scale))
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
    what = ("the whole workload" if scale == 1.0 else f"a {scale:.4f} sample (same |R|:|S| and key distribution)")
    return Rk, Sk, f"|R|={mR}, |S|={mS}: {what}"


def cpu_port_throughput(w, n_gpus=1, reps=2, inputs=None):
    """The oracle's multithreaded host radix join (oracle/oracle_join.c orc_join_check: two radix passes
    through software write-combining buffers + per-partition chained hash join) on ALL host threads --
    the count is passed explicitly, torch.distributed.run exports OMP_NUM_THREADS=1."""
    from oracle import oracle
    Rk, Sk, what = inputs if inputs is not None else cpu_join_sample(w, n_gpus)
    threads = host_threads()
    Rp, Sp = oracle.payload_of_keys(Rk, PAY_SEED_R), oracle.payload_of_keys(Sk, PAY_SEED_S)
    times = []
    for _ in range(reps):
        res, secs = oracle.join_check(Rk, Rp, Sk, Sp, threads, with_time=True)
        times.append(secs)
    best = min(times)
    return {"value": (Rk.size + Sk.size) / best, "unit": UNIT, "cores": threads, "kind": "port",
            "sample": f"{what}; best of {reps}: {best * 1e3:.1f} ms (oracle/oracle_join.c orc_join_check, OpenMP, "
                      f"{threads} threads, SWWC radix passes)",
            "matches": int(res.matches), "checksum": int(res.checksum), "seconds": times}


def reference_arm(args):
    """--impl reference: the CPU port on all host threads (rank 0 only under torchrun).  One step = one
    whole join of cpu_join_sample() -- for workload B the full 128 M x 128 M."""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    w = args.workload
    n_gpus = max(args.gpus, int(os.environ.get("WORLD_SIZE", "1")))
    from oracle import oracle
    oracle.lib()
    inputs = cpu_join_sample(w, n_gpus)
    Rk, Sk, what = inputs
    threads = host_threads()
    Rp, Sp = oracle.payload_of_keys(Rk, PAY_SEED_R), oracle.payload_of_keys(Sk, PAY_SEED_S)
    secs, t_begin = [], time.perf_counter()
    for i in range(args.warmup + args.steps):
        res, s_ = oracle.join_check(Rk, Rp, Sk, Sp, threads, with_time=True)
        if i >= args.warmup:
            secs.append(s_)
        if len(secs) >= 3 and time.perf_counter() - t_begin > 200:      # keep the run within minutes
            break
    per_step = statistics.median(secs)
    v = (Rk.size + Sk.size) / per_step
    line = base_line(args, w, n_gpus)
    cb = {"value": v, "unit": UNIT, "cores": threads, "kind": "port",
          "sample": f"{what}; median of {len(secs)} timed joins: {per_step * 1e3:.1f} ms "
                    f"(oracle/oracle_join.c orc_join_check, OpenMP, {threads} threads, SWWC radix passes)"}
    line.update({"impl": "reference", "value": v, "steps": len(secs), "ms_per_step": per_step * 1e3, "cpu_baseline": cb,
                 "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                 "gpu_launches": 0, "checked": f"matches={int(res.matches)} checksum={int(res.checksum)}",
                 "note": "the reference has no CPU join (its joinCpu is dead code, hash_join_clustered_probe.cu:2013-2059); "
                         "this arm is the oracle's host radix join (structure of partition-primitives.cu:40-125). "
                         "The reference's CUDA kernels are timed by --impl reference-cuda / key reference_cuda."})
    print(json.dumps(line))


CFG5_CACHE = os.path.join(tempfile.gettempdir(), "gj_cfg5_1gpu.json")
CFG5_1GPU_MEASURED = {"value": 75.28e9, "ms_per_step": 53.1,
                      "source": "profiles/r2_final/bench.log, key config5 (round 2, same kernels; no 1-GPU run of this bench found on this box)"}


def run_cfg5_share(torch, gj, n_gpus, rank, dev_index, sharded=None, steps=5, warm=2, opts=()):
    """BASELINE config 5: |R|=|S|=2e9 in total, this GPU generates and joins rows [rank, rank+1) * 2e9/N.
    Returns (ms per step on this rank, matches, checksum of the last step, expected matches, expected checksum)."""
    n_tot = WORKLOADS["cfg5"][0]
    n = n_tot // n_gpus
    dev = torch.device("cuda", dev_index)
    if sharded is None:
        eng = gj.JoinEngine(n, n, dev_index)
    else:
        eng = sharded.ops.engine
    for kv in opts:
        k, v = kv.split("=")
        eng.set_option(k, int(v))
    cols = [torch.empty(n, dtype=torch.int32, device=dev) for _ in range(4)]
    eng.generate_unique(cols[0], cols[1], rank * n, n_tot, 4, PAY_SEED_R)
    eng.generate_unique(cols[2], cols[3], rank * n, n_tot, 5, PAY_SEED_S)
    torch.cuda.synchronize()
    # expected: every key of [0, n_tot) once on each side -> this rank's PROBE shard contributes its own keys
    want = closed_form(torch, cols[2], n_tot)

    def step():
        if sharded is None:
            r = eng.join_aggregate(*cols)
            return r.matches, r.checksum, r.timings.as_dict()
        r = sharded.join_aggregate(*cols, n * n_gpus, n * n_gpus)
        return r.matches, r.checksum, r.phases_ms
    for _ in range(warm):
        step()
    torch.cuda.synchronize()
    if sharded is not None:
        import torch.distributed as dist
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        m, c, tm = step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    if sharded is None:
        eng.close()
    del cols
    torch.cuda.empty_cache()
    return ms, m, c, want, tm


def cfg5_record(ms_step, n_gpus, shuffle=None, tm=None, steps=5, warm=2):
    n_tot = WORKLOADS["cfg5"][0]
    rec = {"workload": workload_name("cfg5"), "scaling": "strong", "n_gpus": n_gpus, "per_gpu_R": n_tot // n_gpus,
           "per_gpu_S": n_tot // n_gpus, "steps": steps, "warmup": warm, "ms_per_step": ms_step,
           "value": 2 * n_tot / (ms_step * 1e-3), "unit": UNIT}
    if shuffle:
        rec["shuffle"] = shuffle
    if n_gpus == 1:
        try:
            with open(CFG5_CACHE, "w") as f:
                json.dump({"value": rec["value"], "ms_per_step": ms_step, "when": time.time()}, f)
        except OSError:
            pass
        rec["speedup_vs_1gpu"] = 1.0
    else:
        base, src = None, None
        try:
            with open(CFG5_CACHE) as f:
                d = json.load(f)
            if time.time() - d["when"] < 6 * 3600:
                base, src = d["value"], f"1-GPU run of this bench on this box ({CFG5_CACHE})"
        except Exception:
            pass
        if base is None:
            base, src = CFG5_1GPU_MEASURED["value"], CFG5_1GPU_MEASURED["source"]
        rec["speedup_vs_1gpu"] = rec["value"] / base
        rec["one_gpu_value"] = base
        rec["one_gpu_source"] = src
    if tm:
        rec["phases_ms_rank0"] = {k: v for k, v in tm.items() if isinstance(v, (int, float))}
    return rec


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
    R, S, expect_m = make_host_keys(gj, w, pins[0][1], pins[2][1])
    hRk, hRp, hSk, hSp = (p[0] for p in pins)

    eng = gj.JoinEngine(nR, nS, 0)
    for kv in args.opt:
        k, v = kv.split("=")
        eng.set_option(k, int(v))
    stream = torch.cuda.Stream()
    eng.use_torch_stream(stream)
    dRk, dSk = hRk.cuda(), hSk.cuda()
    dRp, dSp = torch.empty_like(dRk), torch.empty_like(dSk)
    # real payloads (a function of the key), so that the checksum checks the key/payload pairing of
    # every tuple through both radix passes and the join
    fill_payloads(torch, dRk, dRp, PAY_SEED_R)
    fill_payloads(torch, dSk, dSp, PAY_SEED_S)
    hRp.copy_(dRp); hSp.copy_(dSp)
    want_m, want_c = closed_form(torch, dSk, nR)
    if want_m != expect_m:
        raise SystemExit(f"closed form disagrees with the generator's definition: {want_m} != {expect_m}")
    torch.cuda.synchronize()

    def check(res):
        if res.matches != want_m or res.checksum != want_c:
            raise SystemExit(f"WRONG RESULT: matches={res.matches} checksum={res.checksum}, expected {want_m} / {want_c}")

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
        check(res)
        tms.append(res.timings.as_dict())
    e1.record(stream)
    torch.cuda.synchronize()
    launches = gj.kernel_launch_count() - launches0
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
    med = lambda k: statistics.median(t[k] for t in tms)  # noqa: E731
    roofline = {"bound": "hbm", "peak": peak, "unit": "GB/s", "peak_source": peak_src,
                "per_phase_ms": {"hist_scan_plan": med("hist_ms"), "scatter_passes": med("part_ms"), "join": med("join_ms"),
                                 "total": med("total_ms")}}
    if per_launch:
        achieved = sum(b for b, _ in per_launch) / sum(ms for _, ms in per_launch) / 1e6   # GB/s
        traffic, traffic_src = None, None
        tf = os.path.join(ROOT, "profiles", "scatter_dram_bytes.json")
        if os.path.exists(tf) and w == "B":
            try:
                td = json.load(open(tf))
                traffic, traffic_src = td.get("bytes_per_launch"), f"profiles/scatter_dram_bytes.json ({td.get('source', 'ncu --set full')})"
            except Exception:
                pass
        roofline.update({"kernel": "scatter_kernel (radix scatter pass; %d launches per step)" % sum(1 for x in tms[0]["pass_ms"] if x > 0),
                         "achieved": achieved, "frac": achieved / peak, "traffic": traffic, "traffic_source": traffic_src,
                         "algorithmic_bytes_per_launch": statistics.mean(b for b, _ in per_launch),
                         "avg_launch_ms": statistics.mean(ms for _, ms in per_launch),
                         "pipeline_frac": (44.0 if tms[0]["pass2_bits"] else 28.0) * (nR + nS) / (med("total_ms") * 1e-3) / 1e9 / peak})
    else:   # non-partitioned path (small inputs): build + probe kernels, random access into an L2-resident table
        roofline.update({"kernel": "np_build_kernel + np_probe_kernel (non-partitioned path, L2-resident table)",
                         "achieved": 8.0 * (nR + nS) / (med("total_ms") * 1e-3) / 1e9, "traffic": None,
                         "algorithmic_bytes_per_launch": 8.0 * (nR + nS)})
        roofline["frac"] = roofline["achieved"] / peak

    # ---- end to end: pinned host columns in, result out ----
    for _ in range(max(1, min(args.warmup, 2))):
        check(eng.join_aggregate_host(hRk, hRp, hSk, hSp))
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    e2e_t = []
    for _ in range(args.steps):
        r2 = eng.join_aggregate_host(hRk, hRp, hSk, hSp)
        check(r2)
        e2e_t.append(r2.timings.as_dict())
    torch.cuda.synchronize()
    e2e_ms = (time.perf_counter() - t0) * 1e3 / args.steps
    clocks = sampler.stop()      # sampled across both timed regions (device-resident and end-to-end)
    e2e = {"value": (nR + nS) / (e2e_ms * 1e-3), "unit": UNIT, "ms_per_step": e2e_ms,
           "h2d_bytes_per_step": 8 * (nR + nS), "d2h_bytes_per_step": 32,
           "h2d_ms": statistics.median(t["h2d_ms"] for t in e2e_t),
           "api": "gj_join_aggregate_host (JoinEngine.join_aggregate_host), pinned host columns"}

    # ---- the materialising join on the same inputs: exact-size pair output ----
    mat = None
    if not args.no_materialize:
        out_r = torch.empty(want_m, dtype=torch.int32, device="cuda")
        out_s = torch.empty(want_m, dtype=torch.int32, device="cuda")
        msteps = max(3, min(args.steps, 5))
        mts = []
        for i in range(1 + msteps):
            npairs, mr = eng.join_materialize(dRk, dRp, dSk, dSp, out_r, out_s)
            if npairs != want_m or mr.checksum != want_c:
                raise SystemExit(f"WRONG RESULT (materialize): pairs={npairs} checksum={mr.checksum}, expected {want_m} / {want_c}")
            if i:
                mts.append(mr.timings.as_dict())
        # the written pairs reproduce the checksum: sum over output rows of Pr * Ps
        got_c = 0
        for lo in range(0, want_m, 1 << 26):
            got_c = (got_c + int((out_r[lo:lo + (1 << 26)].to(torch.int64) * out_s[lo:lo + (1 << 26)].to(torch.int64)).sum().item())) & 0xFFFFFFFFFFFFFFFF
        if got_c != want_c:
            raise SystemExit(f"WRONG RESULT (materialize): checksum of the written pairs {got_c} != {want_c}")
        tot = statistics.median(t["total_ms"] for t in mts)
        jm = statistics.median(t["join_ms"] for t in mts)
        mat = {"value": (nR + nS) / (tot * 1e-3), "unit": UNIT, "ms_per_step": tot, "pairs": want_m, "join_ms": jm,
               "join_kernel": {"algorithmic_bytes": 8.0 * (nR + nS) + 8.0 * want_m, "achieved_GBs": (8.0 * (nR + nS) + 8.0 * want_m) / (jm * 1e-3) / 1e9,
                               "frac": (8.0 * (nR + nS) + 8.0 * want_m) / (jm * 1e-3) / 1e9 / peak},
               "api": "gj_join_materialize, device-resident, CUDA events inside the call; pairs staged per CTA (warp-aggregated "
                      "reservation), one global reservation per flush; checked: exact pair count, checksum, and the checksum "
                      "recomputed from the written (Pr, Ps) columns"}
        del out_r, out_s
    eng.close()
    del dRk, dRp, dSk, dSp
    torch.cuda.empty_cache()

    line = base_line(args, w, 1)
    line.update({"value": value, "ms_per_step": ms_step, "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks,
                 "roofline": roofline,
                 "plan": {"radix_bits": tms[0]["radix_bits"], "pass1_bits": tms[0]["pass1_bits"], "pass2_bits": tms[0]["pass2_bits"]},
                 "checked": f"every step: matches == {want_m} and checksum == {want_c} (closed form over the generated keys, payload = mix32(key))"})
    if mat:
        line["materialize"] = mat
    if not args.no_cpu_baseline:
        cb = cpu_port_throughput(w)
        cb.pop("seconds", None)
        if cb["sample"].find("the whole workload") >= 0 and kind == "unique":
            # the CPU port joined the same key sets (any permutation of [0, n)): its aggregate must equal ours
            cb["agrees_with_gpu"] = (cb["matches"], cb["checksum"]) == (want_m, want_c)
        line["cpu_baseline"] = cb
    if not args.no_ref_cuda:
        line["reference_cuda"] = run_reference_cuda(gj, w, 1)
    if w == "B" and not args.no_cfg5:
        ms5, m5, c5, want5, tm5 = run_cfg5_share(torch, gj, 1, 0, 0, opts=args.opt)
        if (m5, c5) != want5:
            raise SystemExit(f"WRONG RESULT (config 5): {m5} {c5}, expected {want5}")
        line["config5"] = cfg5_record(ms5, 1, tm=tm5)
        line["config5"]["checked"] = f"matches == {want5[0]} and checksum == {want5[1]}"
    print(json.dumps(line))


def cfg5_single(args):
    """--workload cfg5 on ONE GPU: the strong-scaling denominator as a line of its own."""
    import torch
    import __graft_entry__ as ge
    gj = ge.load_package()
    torch.cuda.set_device(0)
    launches0 = gj.kernel_launch_count()
    ms5, m5, c5, want5, t = run_cfg5_share(torch, gj, 1, 0, 0, steps=min(args.steps, 5), opts=args.opt)
    if (m5, c5) != want5:
        raise SystemExit(f"WRONG RESULT: {m5} {c5}, expected {want5}")
    n = WORKLOADS["cfg5"][0]
    peak, peak_src = measured_peak()
    passes = 3 if t.get("pass3_bits") else (2 if t.get("pass2_bits") else 1)
    # three-pass plan: every pass 16 B/tuple + the third level's sub-histogram (8 B/tuple)
    alg = (16.0 * passes + (8.0 if passes == 3 else 0.0)) * 2 * n
    line = base_line(args, "cfg5", 1)
    line.update({"steps": min(args.steps, 5), "warmup": 2, "value": 2 * n / (ms5 * 1e-3), "ms_per_step": ms5,
                 "e2e": None, "gpu_launches": int(gj.kernel_launch_count() - launches0),
                 "roofline": {"bound": "hbm", "kernel": f"scatter passes ({passes} per relation" + (" + sub-histogram of the third level)" if passes == 3 else ")"),
                              "achieved": alg / (t["part_ms"] * 1e-3) / 1e9, "peak": peak, "unit": "GB/s", "peak_source": peak_src,
                              "traffic": None, "algorithmic_bytes_per_step": alg,
                              "per_phase_ms": {k: t[k] for k in ("hist_ms", "part_ms", "join_ms", "total_ms")}},
                 "plan": {"radix_bits": t["radix_bits"], "pass1_bits": t["pass1_bits"], "pass2_bits": t["pass2_bits"], "pass3_bits": t.get("pass3_bits")},
                 "checked": f"matches == {want5[0]} and checksum == {want5[1]}"})
    line["roofline"]["frac"] = line["roofline"]["achieved"] / peak
    line["config5"] = cfg5_record(ms5, 1, steps=min(args.steps, 5))
    print(json.dumps(line))


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
        args.shuffle = "pcp"        # measured best at 2, 4 and 8 GPUs (profiles/README.md)
    strong = (w == "cfg5")
    if strong:                      # fixed total, per-GPU share shrinks with N
        nR, nS = nR // world, nS // world
    NR, NS = nR * world, nS * world
    if NR != NS:
        raise SystemExit("unique workloads need |R| == |S|")
    stages = tuple(int(x) for x in args.pcp_stages.split(","))
    # engine capacity: the larger of this workload's and config 5's shard (one context serves both)
    n5 = WORKLOADS["cfg5"][0] // world
    with_cfg5 = (w == "B" and not args.no_cfg5)
    capn = max(nR, n5) if with_cfg5 else nR
    sj = gj.distributed.ShardedJoin(capn, capn, device=local, mode=args.shuffle, overlap=not args.no_overlap, pcp_stages=stages,
                                    pcp_peer_hist=args.pcp_peer_hist)
    for kv in args.opt:
        k, v = kv.split("=")
        sj.ops.engine.set_option(k, int(v))
    eng = sj.ops.engine
    dev = torch.device("cuda", local)
    cols = [torch.empty(n, dtype=torch.int32, device=dev) for n in (nR, nR, nS, nS)]
    eng.generate_unique(cols[0], cols[1], rank * nR, NR, 4, PAY_SEED_R)      # payload = mix32(key): the checksum
    eng.generate_unique(cols[2], cols[3], rank * nS, NS, 5, PAY_SEED_S)      # checks the pairing of every tuple
    # closed form from the keys THIS rank generated: every key of [0, NR) exists once in R, so each of
    # this rank's probe keys matches and contributes payload(key, 40) * payload(key, 50)
    mine = closed_form(torch, cols[2], NR)
    tot = torch.tensor([mine[0], mine[1] - (1 << 64) if mine[1] >= (1 << 63) else mine[1]], dtype=torch.int64, device=dev)
    dist.all_reduce(tot)
    want_m, want_c = int(tot[0].item()), int(tot[1].item()) & 0xFFFFFFFFFFFFFFFF
    if want_m != NS:
        raise SystemExit(f"generator: {want_m} matching probe keys, expected {NS}")
    do_e2e = nR + nS <= 600_000_000
    host = [c.cpu().pin_memory() for c in cols] if do_e2e else None
    torch.cuda.synchronize()

    def step():
        r = sj.join_aggregate(*cols, NR, NS)
        if r.matches != want_m or r.checksum != want_c:
            raise SystemExit(f"rank {rank}: WRONG RESULT {r.matches} {r.checksum}, expected {want_m} {want_c}")
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
    phases = []
    for _ in range(args.steps):
        r = step()
        phases.append(r.phases_ms)
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
    local_rs = (r.local_R, r.local_S)
    del cols, host
    torch.cuda.empty_cache()

    cfg5 = None
    if with_cfg5:
        ms5, m5, c5, want5, tm5 = run_cfg5_share(torch, gj, world, rank, local, sharded=sj)
        t5 = torch.tensor([ms5], device=dev)
        dist.all_reduce(t5, op=dist.ReduceOp.MAX)
        w5 = torch.tensor([want5[0], want5[1] - (1 << 64) if want5[1] >= (1 << 63) else want5[1]], dtype=torch.int64, device=dev)
        dist.all_reduce(w5)
        exp5 = (int(w5[0].item()), int(w5[1].item()) & 0xFFFFFFFFFFFFFFFF)
        if (m5, c5) != exp5:
            raise SystemExit(f"rank {rank}: WRONG RESULT (config 5) {m5} {c5}, expected {exp5}")
        cfg5 = cfg5_record(float(t5.item()), world, shuffle=args.shuffle, tm=tm5)
        cfg5["checked"] = f"matches == {exp5[0]} and checksum == {exp5[1]} (closed form, all-reduced)"

    if rank == 0:
        peak, peak_src = measured_peak()
        line = base_line(args, w, world)
        med = lambda k: statistics.median(p[k] for p in phases if isinstance(p.get(k), (int, float)))  # noqa: E731
        tm = r.phases_ms
        sent = int((nR + nS) * (world - 1) / world)
        line.update({"value": (NR + NS) / (ms_step * 1e-3), "ms_per_step": ms_step,
                     "e2e": None if not do_e2e else {"value": (NR + NS) / (float(e2.item()) * 1e-3), "unit": UNIT, "ms_per_step": float(e2.item()),
                             "h2d_bytes_per_step": 8 * (NR + NS), "d2h_bytes_per_step": 32 * world,
                             "api": "ShardedJoin.join_aggregate after per-rank H2D of pinned host shards"},
                     "gpu_launches": int(launches) * world, "clocks": clocks,
                     "roofline": {"bound": "hbm", "peak": peak, "unit": "GB/s", "peak_source": peak_src, "traffic": None},
                     "shuffle": {"mode": args.shuffle, "tuples_per_gpu_out": sent, "nvlink_peak_GBs": 770.0,
                                 "host_ms_rank0": tm.get("host_ms"), "trace_ms_rank0": tm.get("trace_ms")},
                     "checked": f"every step: matches == {want_m} and checksum == {want_c} (closed form over the generated keys, payload = mix32(key), all-reduced)"})
        if args.shuffle == "pcp" and tm.get("part_R_ms"):
            ph = {k: med(k) for k in ("part_R_ms", "copy_R_ms", "recv_R_ms", "part_S_ms", "copy_S_ms", "recv_S_ms", "tail_ms")}
            copy_ms = ph["copy_R_ms"] + ph["copy_S_ms"]
            # dominant HBM phase that runs ALONE: the source pass of the first relation (layout + first radix pass 16 B/tuple;
            # the phase also holds the counting pass over the chunks this GPU keeps, 8 B/tuple over 1/N of them -- not
            # credited, so the fraction is a lower bound)
            line["roofline"].update({"kernel": "pcp source-side radix pass of R (layout + first pass 16 B/tuple [+ own-chunk counts, not credited]), rank 0",
                                     "achieved": 16.0 * nR / (ph["part_R_ms"] * 1e-3) / 1e9, "algorithmic_bytes_per_launch": 16.0 * nR,
                                     "local_phases_ms": ph})
            line["plan"] = {"gpu_bits": world.bit_length() - 1, "local_bits": tm.get("radix_bits"), "source_pass_bits": tm.get("pass1_bits"),
                            "receiver_pass_bits": tm.get("pass2_bits"), "stages_build_probe": list(stages)}
            line["shuffle"].update({"copy_kernels_ms": copy_ms, "nvlink_out_GBs_per_gpu": 8.0 * sent / (copy_ms * 1e-3) / 1e9,
                                    "nvlink_busy_frac_of_step": copy_ms / ms_step,
                                    "hbm_bytes_per_tuple": 4 + 16 + 8.0 / world + 16.0 * (world - 1) / world + 16 + 8,   # coarse hist, source pass, own-chunk counts, copy out + landing, receiver pass, join
                                    "note": "partition-copy-partition, streamed: first radix pass at the source on [gpu | top local bits] (own chunks "
                                            "straight into the receive buffer), whole first-pass partitions bulk-copied in stages (TMA, global->shared->"
                                            "peer global) with a flag store into every peer after each stage; the receiver partitions and joins a stage "
                                            "while later ones are in flight; tail_ms = last probe byte landed -> last join done"})
            line["shuffle"]["nvlink_frac_of_peer_copy_peak"] = line["shuffle"]["nvlink_out_GBs_per_gpu"] / 770.0
        elif args.shuffle == "pp" and tm.get("local_R_ms"):
            line["roofline"].update({"kernel": "pp local phase of R (hist + first radix pass + fine counts), rank 0",
                                     "achieved": 28.0 * nR / (tm["local_R_ms"] * 1e-3) / 1e9, "algorithmic_bytes_per_launch": 28.0 * nR,
                                     "local_phases_ms": {k: tm.get(k) for k in ("local_R_ms", "push_R_ms", "local_S_ms", "push_S_ms", "join_ms")}})
            line["shuffle"]["scatter_kernel_ms"] = tm.get("shuffle_scatter_ms")
        elif tm.get("pass_ms") and any(tm["pass_ms"][2:]):
            ps = [x for x in tm["pass_ms"][2:] if x > 0]
            line["roofline"].update({"kernel": "scatter_kernel (receiver's radix passes over S, rank 0)",
                                     "achieved": 16.0 * local_rs[1] / (sum(ps) / len(ps) * 1e-3) / 1e9,
                                     "algorithmic_bytes_per_launch": 16.0 * local_rs[1], "avg_launch_ms": sum(ps) / len(ps)})
            line["shuffle"]["scatter_kernel_ms"] = tm.get("shuffle_scatter_ms")
        if line["roofline"].get("achieved"):
            line["roofline"]["frac"] = line["roofline"]["achieved"] / peak
        if cfg5:
            line["config5"] = cfg5
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
    ap.add_argument("--pcp-peer-hist", action="store_true", help="pcp: coarse histograms through the peers' control blocks instead of an NCCL all-gather")
    ap.add_argument("--pcp-stages", default="2,4", help="pcp: copy/receive stages of the building and of the probing relation")
    ap.add_argument("--shuffle", default="auto", choices=["auto", "p2p", "nccl", "dma", "pp", "pcp"],
                    help="multi-GPU exchange: pp = partition locally, last radix pass pushes into the peers; p2p = peer-store "
                         "shuffle first, local passes at the receiver; pcp = first radix pass at the source, whole first-pass partitions "
                         "bulk-copied over NVLink, last pass at the receiver; auto = pcp (measured best, profiles/README.md)")
    ap.add_argument("--opt", action="append", default=[], help="engine option name=value (repeatable)")
    ap.add_argument("--no-overlap", action="store_true", help="multi-GPU: shuffle and local passes back to back")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-ref-cuda", action="store_true")
    ap.add_argument("--no-materialize", action="store_true", help="skip the materialising-join sub-record")
    ap.add_argument("--no-cfg5", action="store_true", help="skip the config-5 (2e9 x 2e9, strong scaling) sub-record")
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
