#!/usr/bin/env python
"""bench.py -- headline benchmark of the brute-force nearest-neighbour hot path.

    python bench.py --gpus N --steps K --warmup W            (N > 1: launched under torchrun)
    python bench.py --impl reference ...                     (the reference's V0 on the host cores)
    python bench.py --gpus N --api multi --workload c4       (ONE process, host arrays, nns_b200_search_multi)

A "step" is one pass of the hot path over one batch: keys_init + fused distance/argmin search of
m queries against the HBM-resident tiled-SoA index of n references (+ the NCCL MIN all-reduce of the
packed keys when the references are sharded) + key unpack, called through the C ABI
(libnns_b200.so).  Metric = pair-distance evaluations per second (m*n / t), whole job.

Default workload at every N = BASELINE.json configs[2] ("C3": k=16, m=262,144, n=16,777,216 uniform
fp32 -- the configuration BASELINE names for 1/2/4/8 GPUs; it fits one GPU).  At N > 1 the references
are sharded (contiguous slices of whole blocks) and the per-query packed (dist, idx) keys are merged by
one NCCL MIN all-reduce: STRONG scaling, a real exchange step; `comm_ms` is the time of that collective
on the device.  At N = 1 the line also carries `also`: 1-GPU records of C2 (k=3) and C4 (k=128), each
with its own roofline / e2e / cpu_baseline.  `--workload` / `--shard` / `--strong` select other cases.

One JSON line on stdout (rank 0).  Timing: CUDA events on the launching stream around every step, L2
flushed between steps, max over ranks; clocks sampled with nvidia-smi during the timed region.
`e2e`: the same metric through the host-pointer API with PINNED host buffers (uploads, index build,
search, [collective], download inside the timed region); `e2e_pageable`: the same with pageable
(malloc'd) arrays, which is what the reference's caller passes (main.cu:27-34).  `index_agreement`:
>= 256 sampled queries against the reference's V0 over the full reference set, at every N.
`cpu_baseline`: V0 (oracle/_ref, else the oracle port) under OpenMP on all host cores on that sample.
"""
from __future__ import annotations

import argparse
import importlib.util
import json
import os
import statistics
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "nns-cuda_b200"))

WORKLOADS = {
    # name: (k, m, n, kind, default sharding at N > 1)
    "c1": (3, 1024, 65536, "uniform", "query"),
    "c2": (3, 65536, 4194304, "uniform", "query"),
    "c3": (16, 262144, 16777216, "uniform", "reference"),
    "c4": (128, 1048576, 1048576, "uniform", "query"),
    "c5": (3, 16777216, 16777216, "clustered", "query"),
    # not a BASELINE config: the reference's m = 1 shapes (main.cu:39-42) scaled up -- the low-arithmetic-intensity
    # case the north-star wants reported as achieved HBM GB/s (reference-parallel kernel, 4(k+1)n bytes per call)
    "m1": (16, 1, 16777216, "uniform", "query"),
    # not a BASELINE config: a contraction longer than the 128 columns of C4 (tcgen05 K-loop kernel, tensor_longk.cu)
    "k256": (256, 262144, 1048576, "uniform", "query"),
}
FP32_LANES_PER_SM = 128
AGREEMENT_SAMPLE = 256


def load_datagen():
    """nns_b200/datagen.py by path: pure numpy, and importing the package would map libnns_b200.so into a
    process (the reference arm) that must not touch the product."""
    spec = importlib.util.spec_from_file_location("nns_datagen", os.path.join(ROOT, "nns-cuda_b200", "nns_b200", "datagen.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def host_cores():
    return len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        return json.load(open(path)), "measured"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "sm_max_mhz": 1965.0}, "fallback"


def ncu_traffic(workload, path_name):
    """dram__bytes_read.sum + dram__bytes_write.sum of the dominant kernel, per launch, from the committed
    ncu --set full captures (profiles/r2_traffic.json, else r1): NOT measured in this run -- ncu cannot run
    inside the timed program.  Returns (bytes or None, source)."""
    for f in ("r2_traffic.json", "r1_traffic.json"):
        try:
            t = json.load(open(os.path.join(ROOT, "profiles", f)))
            if f"{workload}:{path_name}" in t:
                return t[f"{workload}:{path_name}"], f"profiles/{f} (committed ncu --set full capture, not this run)"
        except Exception:
            pass
    return None, None


def make_inputs(name, rank=0, queries=None):
    datagen = load_datagen()
    k, m, n, kind, _ = WORKLOADS[name]
    mq = m if queries is None else queries
    if kind == "uniform":
        s = datagen.uniform_points(mq, k, 1000, 16 * rank)  # rank 0 = stream 0
        r = datagen.uniform_points(n, k, 1000, 1)
    else:
        s, r = datagen.clustered_workload(mq, n, k, 1000 + rank)
    return k, mq, n, s, r


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.gpu)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def mark(self):
        """Start of the timed region: only later samples count (nvidia-smi needs a few hundred ms to
        deliver its first line, so the sampler is started before the warm-up steps)."""
        self.first = len(self.lines)

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, pw, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        first = getattr(self, "first", 0)
        window = self.lines[first:]
        note = None
        if len(window) < 2:  # timed region shorter than the sampling period: use the warm-up samples too
            window, note = self.lines[max(0, first - 5):], "timed region shorter than 2 sampling periods: includes warm-up samples"
        for ln in window:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2])); pw.append(float(f[3]))
            except ValueError:
                continue
            for nm, val in zip(names, f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(nm)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        out = {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(mx), "reasons": sorted(reasons),
               "samples": len(sm), "power_w_max": max(pw)}
        if note:
            out["note"] = note
        return out


def cpu_reference_rate(k, n, s_sample, r, steps=1):
    """The reference's V0 under OpenMP (oracle/_ref when present, else the oracle port) on a query sample
    against the full reference set, on ALL host cores (set explicitly: torchrun exports
    OMP_NUM_THREADS=1).  Returns (pairs/s, info, indices)."""
    from oracle import oracle

    oracle.build()
    oracle.set_threads(host_cores())
    use_ref = oracle.ref() is not None
    fn = oracle.ref_v0_omp if use_ref else oracle.v0_omp
    ms = s_sample.shape[0]
    best = None
    idx = None
    threads = 1
    for _ in range(max(1, steps)):
        t0 = time.perf_counter()
        idx, threads = fn(k, ms, n, s_sample, r, 4)
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    return ms * n / best, {"kind": "reference" if use_ref else "port", "cores": int(threads),
                           "seconds": best, "sample_queries": int(ms)}, idx


def agreement_sample(m, count=AGREEMENT_SAMPLE):
    """Seeded query sample for the V0 check: half even, half odd indices, so that on the clustered workload
    (every 2nd query is an exact copy of a reference) the duplicated-point queries are covered."""
    rng = np.random.default_rng(1000)
    count = min(count, m)
    p = rng.permutation(m)
    even, odd = p[p % 2 == 0][: (count + 1) // 2], p[p % 2 == 1][: count // 2]
    return np.sort(np.concatenate([even, odd]))


def run_reference_arm(args):
    """--impl reference: the reference's own CPU implementation of the path (V0, core.cu:23-53, per query
    chunk under OpenMP) on this box's host cores -- all of them, whatever OMP_NUM_THREADS says -- same
    config / metric / unit.  Nothing of the product is imported or mapped here."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    name = args.workload
    k, m, n, kind, shard = WORKLOADS[name]
    ms = min(m, args.cpu_sample)
    _, _, _, s, r = make_inputs(name, 0, queries=ms)
    times = []
    info = None
    for i in range(args.warmup + args.steps):
        _, info, _ = cpu_reference_rate(k, n, s, r, 1)
        if i >= args.warmup:
            times.append(info["seconds"])
    t = sum(times) / len(times)
    value = ms * n / t
    sample = f"{ms} of {m} queries (seeded, first of stream 0) x all {n} references per step"
    line = {
        "impl": "reference", "metric": "pair_dist_evals_per_s", "value": value, "unit": "pairs/s",
        "queries_per_s": value / n, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": t * 1e3, "higher_is_better": True, "scaling": "strong" if shard == "reference" else "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(name, args.gpus, sharding_text(name, args.gpus, args.shard or shard, args.strong),
                                  "flushed (256 MiB write) between timed steps"),
        "cpu_baseline": {"value": value, "unit": "pairs/s", "cores": info["cores"], "kind": info["kind"], "sample": sample},
        "e2e": {"value": value, "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def sharding_text(name, world, shard, strong):
    k, m, n, kind, _ = WORKLOADS[name]
    if world == 1:
        return "single GPU"
    if shard == "reference":
        return f"reference-sharded x{world} + NCCL MIN all-reduce of packed keys"
    return f"query-sharded x{world} (no data-path collective)" + (f", {m // world} of {m} queries per GPU" if strong else "")


def workload_config(name, n_gpus, sharding, l2):
    k, m, n, kind, _ = WORKLOADS[name]
    return {"workload": f"{name.upper()}: k={k}, m={m} queries, n={n} {kind} fp32 reference points "
                        + (f"(BASELINE.json configs[{list(WORKLOADS).index(name)}])" if name.startswith("c") else
                           "(the reference's m = 1 shape, main.cu:39-42, at 16.7 M references)" if name == "m1" else "(not a BASELINE config)"),
            "k": k, "m": m, "n": n, "distribution": kind, "sharding": sharding, "l2": l2, "n_gpus": n_gpus}


def roofline_of(nns_b200, name, k, m, nloc, flags, kern_ms_avg, clocks, world, timers):
    """Roofline of the dominant kernel of the step (the search: >= 98 % of it), DESIGN.md section 5."""
    pk, pk_src = peaks()
    sm_max_mhz = float(pk.get("sm_max_mhz", 1965.0))
    sm_count = nns_b200.device_sms()
    fp32_peak_tflops = 2.0 * FP32_LANES_PER_SM * sm_count * sm_max_mhz * 1e6 / 1e12
    lane_peak = fp32_peak_tflops * 1e12 / 2.0  # FP32 lane-slots per second
    bf16_peak = float(pk.get("bf16_tflops", 1590.0))
    my_pairs = float(m) * float(nloc)
    rate = my_pairs / (kern_ms_avg * 1e-3)
    path = nns_b200.plan(k, m, nloc, flags, sm_count)["path"]
    tstats = nns_b200.tensor_stats() if path == 2 else None
    fell_back = bool(tstats and tstats["overflow"])  # the screen ran out of candidate space: the FP32 kernel did the work
    fp32_src = f"derived 2*128 lanes*{sm_count} SMs (device query)*{sm_max_mhz:.0f} MHz ({pk_src} sm_max_mhz; FP32 peak is not in MEASURED_PEAKS.json)"
    if path == 0 or (fell_back and k <= 32):
        if flags & nns_b200.FLAG_V0_ROUNDING:
            slots, form = 3 * k, "exact form, V0 rounding"
        elif flags & nns_b200.FLAG_EXACT_FORM:
            slots, form = 2 * k, "exact form (FADD2 + FFMA2 per dimension)"
        else:
            slots, form = k, "norm-expansion screen (FFMA2 per dimension) + exact evaluation of survivors"
        if fell_back:
            form += ("; the tcgen05 screen planned for this shape overflowed its candidate buffer on this data "
                     "and handed over on the device -- the time includes the aborted screen")
        achieved = rate * slots * 2.0 / 1e12
        roofline = {"bound": "fp32", "achieved": achieved, "peak": fp32_peak_tflops, "unit": "TFLOP/s",
                    "frac": achieved / fp32_peak_tflops, "kernel": form, "executed_lane_slots_per_pair": slots,
                    "v0_form_frac": rate * 2.0 * k / lane_peak, "peak_source": fp32_src}
    elif path == 2 and k <= 32:
        # Split-precision tcgen05 screen for low k.  Two resources can bind it: the ALU pipe
        # (ONE minimum per pair in the TMEM epilogue: priced by executed lane-slots) and the tensor
        # pipe (algorithmic 2k FLOPs per pair; the hi/lo column triples and padding are real MMA work but
        # count as zero -- `tensor_frac_executed` shows them).  `bound` = the one with the higher utilisation.
        kp = (tstats or {}).get("kp") or nns_b200.tensor_kp(k)  # the precision mode the index chose (split BF16 or plain F16 columns)
        f16 = (tstats or {}).get("mode") == "f16"
        # FP32 accumulators: FMNMX3 retires two values per instruction at half the FP32 lane rate = 1 lane-slot per pair;
        # F16 accumulators: HMNMX2 / VHMNMX retire two (four) values at the full (half) rate = 0.5 lane-slots per pair
        slots = 0.5 if f16 else 1.0
        alu_frac = rate * slots / lane_peak
        tensor_frac = rate * 2.0 * k / 1e12 / bf16_peak
        tensor_exec = rate * 2.0 * kp / 1e12 / bf16_peak
        kernel = "tcgen05 %s screen (K = %d columns for k = %d) + query image + exact FP32 re-score" % (
            "plain F16 (F16 accumulators)" if f16 else ("split-precision BF16" if kp >= 3 * k else "plain BF16"), kp, k)
        if alu_frac >= tensor_exec:
            roofline = {"bound": "fp32", "achieved": rate * 2.0 * slots / 1e12, "peak": fp32_peak_tflops, "unit": "TFLOP/s", "frac": alu_frac,
                        "kernel": kernel + "; bound by %s lane-slot per pair in the TMEM epilogue" % ("half an HMNMX2" if f16 else "one FMNMX3"),
                        "executed_lane_slots_per_pair": slots, "peak_source": fp32_src + "; the ALU-pipe lane rate equals the FP32 lane rate"}
        else:
            roofline = {"bound": "tensor", "achieved": rate * 2.0 * k / 1e12, "peak": bf16_peak, "unit": "TFLOP/s", "frac": tensor_frac,
                        "kernel": kernel + "; bound by the tensor pipe, which executes %.1fx the algorithmic FLOPs" % (kp / k),
                        "peak_source": f"{pk_src} bf16_tflops (burst, cuBLAS 8192^3)"}
        roofline.update({"alu_min_frac": alu_frac, "tensor_frac": tensor_frac, "tensor_frac_executed": tensor_exec,
                         "v0_form_frac": rate * 2.0 * k / lane_peak})
    elif path == 2:
        achieved = rate * 2.0 * k / 1e12
        roofline = {"bound": "tensor", "achieved": achieved, "peak": bf16_peak, "unit": "TFLOP/s", "frac": achieved / bf16_peak,
                    "peak_source": f"{pk_src} bf16_tflops (burst, cuBLAS 8192^3)",
                    "kernel": "tcgen05 %s screen (K = %d columns) + query image + exact FP32 re-score" % (
                        "F16 (F16 accumulators)" if (tstats or {}).get("mode") == "f16" else "BF16", nns_b200.tensor_kp(k))}
        if pk.get("bf16_tflops_sustained"):
            roofline["frac_of_sustained_peak"] = achieved / float(pk["bf16_tflops_sustained"])
    else:
        hbm = float(pk.get("hbm_gbs", 6650.0))
        if m < 16:
            bytes_per_launch = 4.0 * (k + 1) * nloc + 4.0 * k * m + 8.0 * m
            achieved = bytes_per_launch / (kern_ms_avg * 1e-3) / 1e9
            roofline = {"bound": "hbm", "achieved": achieved, "peak": hbm, "unit": "GB/s", "frac": achieved / hbm,
                        "peak_source": f"{pk_src} hbm_gbs", "kernel": "wide (reference-parallel) kernel"}
        else:
            achieved = rate * 4.0 * k / 1e12
            roofline = {"bound": "fp32", "achieved": achieved, "peak": fp32_peak_tflops, "unit": "TFLOP/s",
                        "frac": achieved / fp32_peak_tflops, "kernel": "wide (reference-parallel) kernel",
                        "executed_lane_slots_per_pair": 2 * k, "peak_source": fp32_src}
    roofline.update({"kernel_ms": kern_ms_avg, "kernel_pairs_per_s": rate})
    if tstats:
        roofline["tensor_stats"] = tstats
    if clocks.get("sm_mhz") and roofline["bound"] != "hbm":
        roofline["frac_at_sampled_clock"] = roofline["frac"] * sm_max_mhz / clocks["sm_mhz"]
    pname = {0: "lowk", 1: "wide", 2: "tensor"}[0 if (fell_back and k <= 32) else path]
    if flags & (nns_b200.FLAG_EXACT_FORM | nns_b200.FLAG_V0_ROUNDING):
        pname += "_exact"
    roofline["traffic"], src = ncu_traffic(name, pname)
    if src:
        roofline["traffic_source"] = src
    # the FP32 kernel north_star describes (V0's formulation, 2k FP32 lane-slots per pair), timed beside the planner's choice
    if k <= 32 and m >= 16 and world == 1 and not (flags & (nns_b200.FLAG_EXACT_FORM | nns_b200.FLAG_V0_ROUNDING)) and timers.get("exact_form"):
        x_ms = timers["exact_form"]()
        x_rate = my_pairs / (x_ms * 1e-3)
        roofline["exact_form_kernel"] = {"kernel_ms": x_ms, "kernel_pairs_per_s": x_rate, "executed_lane_slots_per_pair": 2 * k,
                                         "frac": x_rate * 2.0 * k / lane_peak}
    return roofline


def measure(name, args, env, steps, warmup, full=True):
    """One workload on this rank set: device-resident timing, roofline, agreement with V0, e2e."""
    import torch
    import torch.distributed as dist

    import nns_b200
    from nns_b200 import sharding

    world, rank, device = env["world"], env["rank"], env["device"]
    k, m, n, kind, default_shard = WORKLOADS[name]
    shard = args.shard or default_shard
    if world == 1:
        shard = "query"
    strong_q = args.strong and shard == "query" and world > 1
    weak_q = shard == "query" and world > 1 and not strong_q

    # ---- inputs: synthetic, generated on the host, resident in HBM before the timed region ----
    _, _, _, s_host, r_host = make_inputs(name, rank if weak_q else 0)
    m_total = m
    if strong_q:
        q0, q1 = sharding.query_shard(m, world, rank)
        s_host = np.ascontiguousarray(s_host[q0:q1])
        m = q1 - q0
    r0, r1 = sharding.reference_shard(n, world, rank) if shard == "reference" else (0, n)
    d_q = torch.from_numpy(s_host).to(device)
    d_r = torch.from_numpy(r_host[r0:r1]).to(device)
    index = nns_b200.DeviceIndex(d_r, index_base=r0)
    del d_r
    keys = index.new_keys(m)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=device)  # > 126 MB L2
    stream = torch.cuda.current_stream()
    reduce_keys = world > 1 and shard == "reference"

    def step(evs=None):
        if evs:
            evs[0].record(stream)
        nns_b200._check(nns_b200.lib.nns_b200_keys_init(keys.data_ptr(), m, stream.cuda_stream))
        if evs:
            evs[1].record(stream)
        index.search_keys(d_q, keys, args.flags, stream)
        if evs:
            evs[2].record(stream)
        if reduce_keys:
            dist.all_reduce(keys, op=dist.ReduceOp.MIN)  # packed (dist, idx) keys: exact lowest-index merge
        if evs:
            evs[3].record(stream)
        out = nns_b200.unpack_keys(keys, m, stream)
        if evs:
            evs[4].record(stream)
        return out

    sampler = ClockSampler(env["local_rank"])
    sampler.start()
    for _ in range(warmup):
        flush.zero_()
        idx = step()
    torch.cuda.synchronize()
    ev = [[torch.cuda.Event(enable_timing=True) for _ in range(5)] for _ in range(steps)]
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    launches0 = nns_b200.launch_count()
    sampler.mark()
    wall0 = time.perf_counter()
    for i in range(steps):
        flush.zero_()  # L2 flush between timed iterations (outside the event brackets)
        idx = step(ev[i])
    torch.cuda.synchronize()
    wall = time.perf_counter() - wall0
    clocks = sampler.stop()
    launches = nns_b200.launch_count() - launches0 + (steps if reduce_keys else 0)
    if world > 1:
        dist.barrier()
    t = torch.tensor([sum(e[0].elapsed_time(e[4]) for e in ev), sum(e[1].elapsed_time(e[2]) for e in ev),
                      sum(e[2].elapsed_time(e[3]) for e in ev), wall * 1e3], dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms, kern_total_ms, comm_total_ms, wall_ms = (float(x) for x in t.cpu())
    ms_per_step = total_ms / steps
    kern_ms_avg = kern_total_ms / steps

    job_queries = m_total * (world if weak_q else 1)
    pairs_per_step = float(job_queries) * float(n)
    value = pairs_per_step / (ms_per_step * 1e-3)

    def exact_form_ms():
        reps = 3 if float(m) * float(r1 - r0) < 1e12 else 1  # C3: 4.2 s per pass
        xe = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(reps)]
        for i in range(reps + 1):
            nns_b200._check(nns_b200.lib.nns_b200_keys_init(keys.data_ptr(), m, stream.cuda_stream))
            if i > 0:
                xe[i - 1][0].record(stream)
            index.search_keys(d_q, keys, nns_b200.FLAG_EXACT_FORM, stream)
            if i > 0:
                xe[i - 1][1].record(stream)
        torch.cuda.synchronize()
        return sum(a.elapsed_time(b) for a, b in xe) / len(xe)

    roofline = roofline_of(nns_b200, name, k, m, r1 - r0, args.flags, kern_ms_avg, clocks, world,
                           {"exact_form": exact_form_ms if full and k * m * float(r1 - r0) < 3e14 else None})
    rec = {
        "metric": "pair_dist_evals_per_s", "value": value, "unit": "pairs/s", "queries_per_s": job_queries / (ms_per_step * 1e-3),
        "n_gpus": world, "steps": steps, "warmup": warmup, "ms_per_step": ms_per_step,
        "wall_ms_per_step": wall_ms / steps, "higher_is_better": True,
        "scaling": "weak" if weak_q else "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(name, world, sharding_text(name, world, shard, strong_q), "flushed (256 MiB write) between timed steps"),
        "roofline": roofline, "clocks": clocks, "gpu_launches": int(launches),
        "plan": nns_b200.plan(k, m, r1 - r0, args.flags, nns_b200.device_sms()),
        # the answer is decided by FP32 distances in V0's form; on the tcgen05 path a 16-bit screen (BF16 or F16 operands,
        # see roofline.tensor_stats.mode) only selects which pairs get that exact evaluation
        "dtype_note": "indices from exact FP32 distances; tensor-core screen in " + str((roofline.get("tensor_stats") or {}).get("mode", "n/a")),
    }
    if reduce_keys:
        rec["comm_ms"] = comm_total_ms / steps
        rec["comm"] = f"NCCL all_reduce(MIN) of {m} int64 packed (dist, idx) keys = {8 * m} bytes per step, timed with CUDA events on the launching stream (max over ranks)"
    idx_dev = idx

    # ---- agreement with the reference's V0 on a query sample over the FULL reference set (rank 0, every N) ----
    sample = agreement_sample(m)
    if not args.no_cpu_baseline and rank == 0:
        rate, info, v_idx = cpu_reference_rate(k, n, np.ascontiguousarray(s_host[sample]), r_host, 1)
        g_idx = idx_dev.cpu().numpy()[sample]
        from oracle import oracle
        rep = oracle.check_tie_rule(k, len(sample), n, s_host[sample], r_host, g_idx, v_idx, 1e-5)
        rec["index_agreement"] = {"sampled_queries": int(len(sample)), "identical_to_v0": float((v_idx == g_idx).mean()),
                                  "tie_rule_violations": int(rep["violations"]), "near_tie_accepted": int(rep["near_tie_accepted"]),
                                  "accepted": float(1.0 - rep["violations"] / max(1, len(sample))),
                                  "rule": "north_star: FP64 distance within 1e-5 relative of the minimum; exact ties -> lowest index"}
        rec["cpu_baseline"] = {"value": rate, "unit": "pairs/s", "cores": info["cores"], "kind": info["kind"],
                               "sample": f"{len(sample)} seeded queries of {m} x all {n} references, {info['seconds']:.2f} s"}

    # ---- end to end through the host-pointer API ----
    if not args.no_e2e:
        ne2e = max(2, min(steps, 5))
        out = np.empty(m, dtype=np.int32)

        def e2e_single(sp, rp):
            nns_b200.search_host(k, m, n, sp, rp, out)
            return out

        def e2e_sharded(s_pin, r_pin):
            # reference-sharded across ranks: upload the slice + queries, build, search, MIN all-reduce, unpack, download
            dq = s_pin.to(device, non_blocking=True)
            dr = r_pin.to(device, non_blocking=True)
            ix = nns_b200.DeviceIndex(dr, index_base=r0)
            kk = ix.new_keys(m)
            ix.search_keys(dq, kk, args.flags, stream)
            dist.all_reduce(kk, op=dist.ReduceOp.MIN)
            return nns_b200.unpack_keys(kk, m, stream).cpu().numpy()

        def e2e_gathered(s_pin, r_pin):
            # query-sharded across ranks: every rank uploads ONE slice of the references, builds it, and the built
            # slices are all-gathered over NVLink (sharding.gather_built_index) -- not `world` full uploads
            dq = s_pin.to(device, non_blocking=True)
            ix = sharding.gather_built_index(k, n, r_pin, rank, world, device, stream)
            return ix.search(dq, args.flags, stream).cpu().numpy()

        def timed(fn, a, b):
            for _ in range(2):
                res = fn(a, b)
            if world > 1:
                dist.barrier()
            t0 = time.perf_counter()
            for _ in range(ne2e):
                res = fn(a, b)
            dt = (time.perf_counter() - t0) / ne2e
            te = torch.tensor([dt], dtype=torch.float64, device=device)
            if world > 1:
                dist.all_reduce(te, op=dist.ReduceOp.MAX)
            return float(te.item()), res

        r_mine = r_host[r0:r1] if shard == "reference" else r_host
        s_pin = torch.from_numpy(s_host).pin_memory()
        r_pin = torch.from_numpy(r_mine).pin_memory()
        if world == 1:
            e2e_s, res = timed(e2e_single, s_pin.data_ptr(), r_pin.data_ptr())
            api = "nns_b200_search_host (pinned host buffers; H2D + index build + search + D2H)"
            h2d = int(s_pin.numel() * 4 + r_pin.numel() * 4)
        elif shard == "reference":
            e2e_s, res = timed(e2e_sharded, s_pin, r_pin)
            api = "per rank: pinned H2D of its reference slice + all queries, nns_b200_index_build, search_keys, NCCL MIN all-reduce, keys_unpack, D2H"
            h2d = int(s_pin.numel() * 4 + r_pin.numel() * 4)
        else:
            e2e_s, res = timed(e2e_gathered, s_pin, r_pin)
            api = ("per rank: pinned H2D of its queries + ONE 1/N slice of the references, nns_b200_index_build_part, NCCL all-gather of the built "
                   "index over NVLink, search, D2H")
            h2d = int(s_pin.numel() * 4 + (r_pin.numel() * 4) // world)
        rec["e2e"] = {"value": pairs_per_step / e2e_s, "unit": "pairs/s", "ms_per_step": e2e_s * 1e3,
                      "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": int(m * 4), "steps": ne2e, "api": api}
        same = np.array_equal(res, idx_dev.cpu().numpy())
        rec["e2e"]["result_equals_device_path"] = bool(same)
        if world == 1:
            # the reference's caller passes malloc'd arrays (main.cu:27-34): the same call with pageable memory
            e2e_p, res_p = timed(e2e_single, s_host, r_host)
            rec["e2e_pageable"] = {"value": pairs_per_step / e2e_p, "unit": "pairs/s", "ms_per_step": e2e_p * 1e3,
                                   "api": "nns_b200_search_host (pageable numpy arrays through the pinned staging ring)",
                                   "vs_pinned": e2e_p / e2e_s, "result_equals_device_path": bool(np.array_equal(res_p, idx_dev.cpu().numpy()))}
        del s_pin, r_pin
    del index, keys, d_q, flush
    torch.cuda.empty_cache()
    return rec


def ref_gpu_record():
    """The reference's best valid GPU kernel on this box beside this engine, on the reference's own data and
    timing (C1; V7 core.cu:589-633 through oracle/_ref/ref_gpu_probe; wall clock around the whole callback like
    main.cu:73-76).  None when the probe binary is not present."""
    probe = os.path.join(ROOT, "oracle", "_ref", "ref_gpu_probe")
    if not os.path.exists(probe):
        return None
    import nns_b200
    from nns_b200 import datagen
    from oracle import oracle

    k, m, n = 3, 1024, 65536
    rec = {"workload": "C1 k=3 m=1024 n=65536, the reference's generator (srand(1000), main.cu:24-35)",
           "timing": "wall clock around the whole host-pointer callback (malloc/H2D/kernel/D2H inside), best of 20 after 1 warm-up"}
    s, r = datagen.reference_rand_sample(k, m, n, 1000)
    v = oracle.ref_v0(k, m, n, s, r) if oracle.ref() is not None else oracle.v0(k, m, n, s, r)
    try:
        for var in (7, 9):
            with tempfile.NamedTemporaryFile(suffix=".bin") as f:
                outp = subprocess.run([probe, str(var), str(k), str(m), str(n), "20", f.name], capture_output=True, text=True, timeout=120)
                line = [ln for ln in outp.stdout.splitlines() if ln.startswith("ref_gpu")]
                if outp.returncode != 0 or not line:
                    rec[f"v{var}"] = {"error": (outp.stderr or outp.stdout)[-200:]}
                    continue
                kv = dict(x.split("=") for x in line[0].split()[1:])
                g = np.fromfile(f.name, dtype=np.int32)
                rec[f"v{var}"] = {"best_ms": float(kv["best_ms"]), "median_ms": float(kv["median_ms"]),
                                  "identical_to_v0": float((g == v).mean()) if g.size == m else None}
    except Exception as e:  # the probe is a side record; never fail the bench over it
        rec["error"] = str(e)[:200]
    ts = []
    for _ in range(21):
        t0 = time.perf_counter()
        g = nns_b200.cudaCall(k, m, n, s, r)
        ts.append((time.perf_counter() - t0) * 1e3)
    rec["this_engine"] = {"best_ms": min(ts[1:]), "median_ms": statistics.median(ts[1:]), "identical_to_v0": float((g == v).mean()),
                          "api": "nns_b200_cudaCall (drop-in symbol, pageable arrays)"}
    return rec


def tree_record():
    """Side record (not the metric): the exact KD-tree path (csrc/kdtree.cu, SURVEY 8f row n4 -- the reference's
    V11 is an empty stub) on C2's shape: build time (upload + Morton sort + leaves + boxes on the GPU), search time
    through nns_b200_tree_search with host
    arrays, and agreement with the brute-force answer of the same library and with V0 on a sample."""
    import nns_b200

    k, m, n = 3, 65536, 4194304
    _, _, _, s, r = make_inputs("c2", 0)
    t0 = time.perf_counter()
    tree = nns_b200.HostTree(k, n, r)
    build_s = time.perf_counter() - t0
    tree.search(1024, s[:1024])
    ts = []
    for _ in range(5):
        t0 = time.perf_counter()
        g = tree.search(m, s)
        ts.append((time.perf_counter() - t0) * 1e3)
    brute = nns_b200.search_host(k, m, n, s, r)
    sample = agreement_sample(m)
    _, _, v_idx = cpu_reference_rate(k, n, np.ascontiguousarray(s[sample]), r, 1)
    tree.close()
    return {"workload": "C2's shape through the exact KD-tree (k=3, m=65536, n=4194304)", "build_ms": build_s * 1e3,
            "build": "nns_b200_tree_create: host array in, upload + GPU Morton-order build (first call: includes one-time allocations)",
            "search_ms": min(ts), "search_ms_median": statistics.median(ts),
            "timing": "wall clock around nns_b200_tree_search (query upload + search + download)",
            "identical_to_brute_force": float((g == brute).mean()), "identical_to_v0_sample": float((g[sample] == v_idx).mean())}


def run_multi_api(args):
    """--api multi: ONE process, host arrays, nns_b200_search_multi on N GPUs -- the reference's V8/V9 shape
    (core.cu:965-1057).  Wall clock around the whole call (everything is inside: uploads over N PCIe links,
    build + NVLink all-gather / key merge, search, download)."""
    import torch

    import nns_b200

    name = args.workload
    k, m, n, kind, default_shard = WORKLOADS[name]
    shard = args.shard or default_shard
    mode = 1 if shard == "reference" else 0
    G = args.gpus
    assert torch.cuda.device_count() >= G, f"--gpus {G} but {torch.cuda.device_count()} visible"
    _, _, _, s, r = make_inputs(name, 0)
    if args.pinned:
        s_t, r_t = torch.from_numpy(s).pin_memory(), torch.from_numpy(r).pin_memory()
        sp, rp = s_t.data_ptr(), r_t.data_ptr()
    else:
        sp, rp = s.ctypes.data, r.ctypes.data
    out = np.empty(m, dtype=np.int32)
    sampler = ClockSampler(0)
    sampler.start()

    def call():
        nns_b200._check(nns_b200.lib.nns_b200_search_multi(k, m, n, sp, rp, out.ctypes.data, G, mode))

    for _ in range(args.warmup):
        call()
    launches0 = nns_b200.launch_count()
    sampler.mark()
    ts = []
    for _ in range(args.steps):
        t0 = time.perf_counter()
        call()
        ts.append(time.perf_counter() - t0)
    clocks = sampler.stop()
    t = sum(ts) / len(ts)
    value = float(m) * float(n) / t
    line = {"metric": "pair_dist_evals_per_s", "value": value, "unit": "pairs/s", "queries_per_s": m / t, "n_gpus": G,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": t * 1e3, "best_ms": min(ts) * 1e3, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(name, G, f"nns_b200_search_multi, one process, {G} GPUs, shard_mode {mode} "
                                      + ("(reference-sharded; packed keys folded into GPU 0 by system-scope red.min over NVLink)" if mode
                                         else "(query-sharded; sharded ingest, build kernels store every slice into all peers' indices)"),
                                      "host arrays re-ingested every step"),
            "timing": "wall clock around the whole C-ABI call (host arrays in, host indices out)", "clocks": clocks,
            "gpu_launches": int(nns_b200.launch_count() - launches0),
            "e2e": {"value": value, "unit": "pairs/s", "ms_per_step": t * 1e3, "h2d_bytes_per_step": int(4 * k * (m * (G if mode else 1) + n)),
                    "d2h_bytes_per_step": int(4 * m), "api": "nns_b200_search_multi (" + ("pinned" if args.pinned else "pageable") + " host arrays)"}}
    if not args.no_cpu_baseline:
        sample = agreement_sample(m)
        rate, info, v_idx = cpu_reference_rate(k, n, np.ascontiguousarray(s[sample]), r, 1)
        from oracle import oracle
        rep = oracle.check_tie_rule(k, len(sample), n, s[sample], r, out[sample], v_idx, 1e-5)
        line["index_agreement"] = {"sampled_queries": int(len(sample)), "identical_to_v0": float((v_idx == out[sample]).mean()),
                                   "tie_rule_violations": int(rep["violations"]), "accepted": float(1.0 - rep["violations"] / len(sample))}
        line["cpu_baseline"] = {"value": rate, "unit": "pairs/s", "cores": info["cores"], "kind": info["kind"],
                                "sample": f"{len(sample)} seeded queries x all {n} references, {info['seconds']:.2f} s"}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--api", default="ranks", choices=["ranks", "multi"],
                    help="ranks: one process per GPU (torchrun at N > 1); multi: one process driving N GPUs through nns_b200_search_multi")
    ap.add_argument("--workload", default="c3", choices=sorted(WORKLOADS))
    ap.add_argument("--shard", default=None, choices=["query", "reference"])
    ap.add_argument("--strong", action="store_true",
                    help="query-sharded: split the workload's m queries across the ranks (the configuration as "
                         "BASELINE.json names it for C4/C5) instead of m queries per rank (weak scaling)")
    ap.add_argument("--flags", type=lambda x: int(x, 0), default=0, help="nns_b200 flags word (tuning overrides)")
    ap.add_argument("--cpu-sample", type=int, default=AGREEMENT_SAMPLE, help="queries per step of the --impl reference arm")
    ap.add_argument("--also", default=None, help="comma-separated 1-GPU side workloads (default at N = 1 on c3: c2,c4; 'none' to skip)")
    ap.add_argument("--pinned", action="store_true", help="--api multi: pin the host arrays")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-ref-gpu", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference_arm(args)
        return
    if args.warmup < 3:
        args.warmup = 3  # timing rules: W >= 3
    if args.api == "multi":
        run_multi_api(args)
        return

    import torch
    import torch.distributed as dist

    import nns_b200

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the CUDA path is the product; there is no CPU fallback to time")
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=device)
    nns_b200.init(local_rank)
    env = {"world": world, "rank": rank, "local_rank": local_rank, "device": device}

    line = measure(args.workload, args, env, args.steps, args.warmup)
    if world == 1:
        also = args.also if args.also is not None else ("c2,c4" if args.workload == "c3" and args.flags == 0 else "none")
        recs = []
        for nm in [x for x in also.split(",") if x and x != "none"]:
            sub = argparse.Namespace(**vars(args))
            sub.shard, sub.strong = None, False
            recs.append(measure(nm, sub, env, max(3, min(args.steps, 10)), 3))
        if recs:
            line["also"] = recs
        if not args.no_ref_gpu and not args.no_cpu_baseline:
            rg = ref_gpu_record()
            if rg:
                line["ref_gpu"] = rg
            if recs:  # the default (driver) run only
                try:
                    line["tree"] = tree_record()
                except Exception as e:  # a side record never fails the bench
                    line["tree"] = {"error": str(e)[:200]}
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
