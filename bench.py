#!/usr/bin/env python
"""bench.py -- headline benchmark of the brute-force nearest-neighbour hot path.

    python bench.py --gpus N --steps K --warmup W            (N > 1: launched under torchrun)
    python bench.py --impl reference ...                     (the reference's V0 on the host cores)

A "step" is one pass of the hot path over one batch: keys_init + fused distance/argmin search of
m queries against the HBM-resident tiled-SoA index of n references + key unpack, called through
the C ABI (libnns_b200.so).  Metric = pair-distance evaluations per second (m*n / t), whole job.
Default workload = BASELINE.json configs[1] ("C2": k=3, m=65,536, n=4,194,304 uniform fp32); the
library's planner picks the kernel (C2: the split-precision tcgen05 screen + exact FP32 re-score),
`--flags` forces another one (2 = FP32 screened kernel, 16 = V0's formulation on the FP32 pipe).
`roofline` describes the kernel that ran, against the resource that binds it.
With N > 1 the queries are sharded (each rank searches its own m queries against the replicated
reference set: weak scaling, no data-path collective); `--shard reference` splits the references
instead and merges the packed (dist, idx) keys with an NCCL MIN all-reduce (BASELINE config C3).

One JSON line on stdout (rank 0).  Timing: CUDA events on the launching stream around every
step, L2 flushed between steps, max over ranks; clocks sampled with nvidia-smi during the timed
region; `e2e` is the same metric through nns_b200_search_host with pinned HOST buffers (H2D of
queries + references, index build, search, D2H of the indices inside the timed region);
`cpu_baseline` is the reference's V0 (oracle/_ref, else the oracle port) under OpenMP on a
bounded query sample.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
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
}
SM_COUNT = 148
FP32_LANES_PER_SM = 128


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        p = json.load(open(path))
        return p, "measured"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "sm_max_mhz": 1965.0}, "fallback"


def ncu_traffic(workload, path_name):
    """dram__bytes_read.sum + dram__bytes_write.sum of the dominant kernel, per launch, from the
    committed ncu --set full captures (profiles/r1_traffic.json); None when no capture matches."""
    try:
        t = json.load(open(os.path.join(ROOT, "profiles", "r1_traffic.json")))
        return t.get(f"{workload}:{path_name}")
    except Exception:
        return None


def make_inputs(name, rank=0):
    from nns_b200 import datagen

    k, m, n, kind, _ = WORKLOADS[name]
    if kind == "uniform":
        s = datagen.uniform_points(m, k, 1000, 16 * rank)  # rank 0 = stream 0
        r = datagen.uniform_points(n, k, 1000, 1)
    else:
        s, r = datagen.clustered_workload(m, n, k, 1000 + rank)
    return k, m, n, s, r


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
    """The reference's V0 under OpenMP (oracle/_ref when present, else the oracle port) on a query
    sample against the full reference set.  Returns (pairs/s, info)."""
    from oracle import oracle

    oracle.build()
    use_ref = oracle.ref() is not None
    fn = oracle.ref_v0_omp if use_ref else oracle.v0_omp
    ms = s_sample.shape[0]
    best = None
    idx = None
    for _ in range(max(1, steps)):
        t0 = time.perf_counter()
        idx, threads = fn(k, ms, n, s_sample, r, 4)
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    return ms * n / best, {"kind": "reference" if use_ref else "port", "cores": int(threads),
                           "seconds": best, "sample_queries": int(ms)}, idx


def run_reference_arm(args):
    """--impl reference: the reference's own CPU implementation of the path (V0, core.cu:23-53,
    per query chunk under OpenMP) on this box's host cores, same config / metric / unit."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    name = args.workload
    k, m, n, kind, _ = WORKLOADS[name]
    from nns_b200 import datagen

    ms = min(m, args.cpu_sample)
    if kind == "uniform":
        r = datagen.uniform_points(n, k, 1000, 1)
        s = datagen.uniform_points(ms, k, 1000, 0)
    else:
        s, r = datagen.clustered_workload(ms, n, k, 1000)
    times = []
    threads = 1
    info = None
    for i in range(args.warmup + args.steps):
        rate, info, _ = cpu_reference_rate(k, n, s, r, 1)
        if i >= args.warmup:
            times.append(info["seconds"])
        threads = info["cores"]
    t = sum(times) / len(times)
    value = ms * n / t
    sample = f"{ms} of {m} queries (seeded, first of stream 0) x all {n} references per step"
    line = {
        "impl": "reference", "metric": "pair_dist_evals_per_s", "value": value, "unit": "pairs/s",
        "queries_per_s": value / n, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": t * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": workload_config(name, args.gpus, "none (host cores only)", "n/a"),
        "cpu_baseline": {"value": value, "unit": "pairs/s", "cores": threads, "kind": info["kind"], "sample": sample},
        "e2e": {"value": value, "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def workload_config(name, n_gpus, sharding, l2):
    k, m, n, kind, _ = WORKLOADS[name]
    return {"workload": f"{name.upper()}: k={k}, m={m} queries, n={n} {kind} fp32 reference points "
                        + (f"(BASELINE.json configs[{list(WORKLOADS).index(name)}])" if name != "m1" else "(the reference's m = 1 shape, main.cu:39-42, at 16.7 M references)"),
            "k": k, "m": m, "n": n, "distribution": kind, "sharding": sharding, "l2": l2,
            "queries_per_gpu": m if sharding.startswith("query") else m, "n_gpus": n_gpus}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS))
    ap.add_argument("--shard", default=None, choices=["query", "reference"])
    ap.add_argument("--strong", action="store_true",
                    help="query-sharded: split the workload's m queries across the ranks (the configuration as "
                         "BASELINE.json names it for C4/C5) instead of m queries per rank (weak scaling)")
    ap.add_argument("--flags", type=lambda x: int(x, 0), default=0, help="nns_b200 flags word (tuning overrides)")
    ap.add_argument("--cpu-sample", type=int, default=2048, help="queries in the CPU-baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "b200":
        args.warmup = max(args.warmup, 3)  # timing rules: W >= 3

    if args.impl == "reference":
        run_reference_arm(args)
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
    n_gpus = world
    name = args.workload
    k, m, n, kind, default_shard = WORKLOADS[name]
    shard = args.shard or default_shard
    if world == 1:
        shard = "query"

    # ---- inputs: synthetic, generated on the host, resident in HBM before the timed region ----
    strong_q = args.strong and shard == "query" and world > 1
    kk, mm, nn, s_host, r_host = make_inputs(name, rank if (shard == "query" and not strong_q) else 0)
    m_total = m
    if strong_q:
        from nns_b200 import sharding

        q0, q1 = sharding.query_shard(m, world, rank)
        s_host = np.ascontiguousarray(s_host[q0:q1])
        m = q1 - q0
    if shard == "reference":
        blocks = (n + 127) // 128
        per = ((blocks + world - 1) // world) * 128
        r0 = min(n, rank * per)
        r1 = min(n, r0 + per)
    else:
        r0, r1 = 0, n
    nns_b200.init(local_rank)
    d_q = torch.from_numpy(s_host).to(device)
    d_r = torch.from_numpy(r_host[r0:r1]).to(device)
    index = nns_b200.DeviceIndex(d_r, index_base=r0)
    del d_r
    keys = index.new_keys(m)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=device)  # > 126 MB L2
    stream = torch.cuda.current_stream()

    def step():
        nns_b200._check(nns_b200.lib.nns_b200_keys_init(keys.data_ptr(), m, stream.cuda_stream))
        index.search_keys(d_q, keys, args.flags, stream)
        if world > 1 and shard == "reference":
            dist.all_reduce(keys, op=dist.ReduceOp.MIN)  # packed (dist, idx) keys: exact lowest-index merge
        return nns_b200.unpack_keys(keys, m, stream)

    sampler = ClockSampler(local_rank)
    sampler.start()
    for _ in range(args.warmup):
        flush.zero_()
        idx = step()
    torch.cuda.synchronize()

    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True),
           torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    launches0 = nns_b200.launch_count()
    sampler.mark()
    wall0 = time.perf_counter()
    for i in range(args.steps):
        flush.zero_()  # L2 flush between timed iterations (outside the event brackets)
        e0, k0, k1, e1 = ev[i]
        e0.record(stream)
        nns_b200._check(nns_b200.lib.nns_b200_keys_init(keys.data_ptr(), m, stream.cuda_stream))
        k0.record(stream)
        index.search_keys(d_q, keys, args.flags, stream)
        k1.record(stream)
        if world > 1 and shard == "reference":
            dist.all_reduce(keys, op=dist.ReduceOp.MIN)
        idx = nns_b200.unpack_keys(keys, m, stream)
        e1.record(stream)
    torch.cuda.synchronize()
    wall = time.perf_counter() - wall0
    clocks = sampler.stop()
    launches = nns_b200.launch_count() - launches0
    if world > 1:
        dist.barrier()
    step_ms = [e0.elapsed_time(e1) for (e0, _, _, e1) in ev]
    kern_ms = [a.elapsed_time(b) for (_, a, b, _) in ev]
    total_ms = sum(step_ms)
    t = torch.tensor([total_ms, sum(kern_ms), wall * 1e3], dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms, kern_total_ms, wall_ms = (float(x) for x in t.cpu())
    ms_per_step = total_ms / args.steps
    kern_ms_avg = kern_total_ms / args.steps

    job_queries = m_total * (world if (shard == "query" and not strong_q) else 1)
    pairs_per_step = float(job_queries) * float(n)
    value = pairs_per_step / (ms_per_step * 1e-3)
    queries_per_s = job_queries / (ms_per_step * 1e-3)

    # ---- roofline of the dominant kernel (lowk/wide search): FP32 pipe, SURVEY 8(d) ----
    pk, pk_src = peaks()
    sm_max_mhz = float(pk.get("sm_max_mhz", 1965.0))
    fp32_peak_tflops = 2.0 * FP32_LANES_PER_SM * SM_COUNT * sm_max_mhz * 1e6 / 1e12
    my_pairs = float(m) * float(r1 - r0)
    kern_pairs_per_s = my_pairs / (kern_ms_avg * 1e-3)
    lane_peak = fp32_peak_tflops * 1e12 / 2.0  # FP32 lane-slots per second
    path = nns_b200.plan(k, m, r1 - r0, args.flags)["path"]
    bf16_peak = float(pk.get("bf16_tflops", 1590.0))

    def exact_form_side_measurement():
        # the same workload on the exact-form FP32 kernel (outside the headline timing): V0's formulation,
        # 2k FP32 lane-slots per pair -- the figure the north-star's 70 % FP32 bar refers to
        xe = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(3)]
        for i in range(4):
            nns_b200._check(nns_b200.lib.nns_b200_keys_init(keys.data_ptr(), m, stream.cuda_stream))
            if i > 0:
                xe[i - 1][0].record(stream)
            index.search_keys(d_q, keys, nns_b200.FLAG_EXACT_FORM, stream)
            if i > 0:
                xe[i - 1][1].record(stream)
        torch.cuda.synchronize()
        x_ms = sum(a.elapsed_time(b) for a, b in xe) / len(xe)
        x_rate = my_pairs / (x_ms * 1e-3)
        return {"kernel_ms": x_ms, "kernel_pairs_per_s": x_rate, "executed_lane_slots_per_pair": 2 * k,
                "frac": x_rate * 2.0 * k / lane_peak}

    tstats = nns_b200.tensor_stats() if path == 2 else None
    fell_back = bool(tstats and tstats["overflow"])  # the screen ran out of candidate space: the FP32 kernel did the work
    if path == 0 or (fell_back and k <= 32):
        # Executed FP32 lane-slots per pair: the screened kernel evaluates
        # s = |r|^2 - 2q.r with k FMAs per pair; the exact-form kernel runs V0's k subtractions +
        # k FMAs (2k; 3k with separately rounded mul/add).  `frac` is computed from the slots the
        # kernel actually executes; `v0_form_frac` prices every pair at V0's 2k slots (the
        # formulation SURVEY.md 8(d) and the north-star's 70 % bar refer to) and can exceed 1.
        if args.flags & nns_b200.FLAG_V0_ROUNDING:
            slots, form = 3 * k, "exact form, V0 rounding"
        elif args.flags & nns_b200.FLAG_EXACT_FORM:
            slots, form = 2 * k, "exact form (FADD2 + FFMA2 per dimension)"
        else:
            slots, form = k, "norm-expansion screen (FFMA2 per dimension) + exact evaluation of survivors"
        if fell_back:
            form += ("; the tcgen05 screen planned for this shape overflowed its candidate buffer on this data "
                     "(dense near-ties) and handed over on the device -- the time includes the aborted screen")
        achieved = kern_pairs_per_s * slots * 2.0 / 1e12
        roofline = {"bound": "fp32", "achieved": achieved, "peak": fp32_peak_tflops, "unit": "TFLOP/s",
                    "frac": achieved / fp32_peak_tflops, "traffic": None,
                    "kernel": form, "executed_lane_slots_per_pair": slots,
                    "v0_form_frac": kern_pairs_per_s * 2.0 * k / lane_peak,
                    "peak_source": f"derived 2*128 lanes*148 SMs*{sm_max_mhz:.0f} MHz ({pk_src} sm_max_mhz; FP32 peak is not in MEASURED_PEAKS.json)",
                    "kernel_ms": kern_ms_avg, "kernel_pairs_per_s": kern_pairs_per_s}
        if tstats:
            roofline["tensor_stats"] = tstats
        if clocks.get("sm_mhz"):
            roofline["frac_at_sampled_clock"] = roofline["frac"] * sm_max_mhz / clocks["sm_mhz"]
        if not (args.flags & (nns_b200.FLAG_EXACT_FORM | nns_b200.FLAG_V0_ROUNDING)) and m >= 16 and world == 1:
            roofline["exact_form_kernel"] = exact_form_side_measurement()
    elif path == 2 and k <= 32:
        # Split-precision tcgen05 screen for low k.  Two resources can bind it:
        #  * the ALU pipe: every pair costs ONE minimum in the epilogue that reduces the TMEM accumulators
        #    (FMNMX3 retires two new values per instruction at half rate: 128 pairs/clk/SM, the same lane
        #    rate as the FP32 pipe).  SURVEY.md 8(d): a kernel that executes fewer than V0's 2k lane-slots
        #    per pair is priced at the slots it executes -> 1 per pair against 128 lanes x 148 SMs x f;
        #  * the tensor pipe, once the contraction is long enough (k = 16: 3k + 3 columns padded to 64):
        #    algorithmic work 2k FLOPs per pair against the measured BF16 peak; the hi/lo column triples
        #    and the padding are real MMA work but count as zero (`tensor_frac_executed` shows them).
        # The line reports the one with the higher utilisation as `bound`, the other beside it.
        kp = nns_b200.tensor_kp(k)
        alu_frac = kern_pairs_per_s * 1.0 / lane_peak
        tensor_frac = kern_pairs_per_s * 2.0 * k / 1e12 / bf16_peak
        tensor_exec = kern_pairs_per_s * 2.0 * kp / 1e12 / bf16_peak
        kernel = ("tcgen05 split-precision BF16 screen (K = %d columns for k = %d) + query image + exact FP32 re-score" % (kp, k))
        if alu_frac >= tensor_exec:
            roofline = {"bound": "fp32", "achieved": kern_pairs_per_s * 2.0 / 1e12, "peak": fp32_peak_tflops, "unit": "TFLOP/s",
                        "frac": alu_frac, "traffic": None,
                        "kernel": kernel + "; bound by one FMNMX3 lane-slot per pair in the TMEM epilogue",
                        "executed_lane_slots_per_pair": 1,
                        "peak_source": f"derived 2*128 lanes*148 SMs*{sm_max_mhz:.0f} MHz ({pk_src} sm_max_mhz; the ALU-pipe lane rate "
                                       f"equals the FP32 lane rate)"}
        else:
            roofline = {"bound": "tensor", "achieved": kern_pairs_per_s * 2.0 * k / 1e12, "peak": bf16_peak, "unit": "TFLOP/s",
                        "frac": tensor_frac, "traffic": None,
                        "kernel": kernel + "; bound by the tensor pipe, which executes %.1fx the algorithmic FLOPs" % (kp / k),
                        "peak_source": f"{pk_src} bf16_tflops (burst, cuBLAS 8192^3)"}
        roofline.update({"alu_min_frac": alu_frac, "tensor_frac": tensor_frac, "tensor_frac_executed": tensor_exec,
                         "v0_form_frac": kern_pairs_per_s * 2.0 * k / lane_peak,
                         "kernel_ms": kern_ms_avg, "kernel_pairs_per_s": kern_pairs_per_s, "tensor_stats": tstats})
        if clocks.get("sm_mhz"):
            roofline["frac_at_sampled_clock"] = roofline["frac"] * sm_max_mhz / clocks["sm_mhz"]
        if world == 1:
            roofline["exact_form_kernel"] = exact_form_side_measurement()
    elif path == 2:
        # tcgen05 path: 2k FLOPs per pair (the -2 q.r contraction only; norms, epilogue and the exact
        # re-score count as zero) against the measured dense BF16 peak
        achieved = kern_pairs_per_s * 2.0 * k / 1e12
        roofline = {"bound": "tensor", "achieved": achieved, "peak": bf16_peak, "unit": "TFLOP/s", "frac": achieved / bf16_peak,
                    "traffic": None, "peak_source": f"{pk_src} bf16_tflops (burst, cuBLAS 8192^3)",
                    "kernel": "tcgen05 BF16 screen (K = %d columns) + query image + exact FP32 re-score" % nns_b200.tensor_kp(k),
                    "kernel_ms": kern_ms_avg, "kernel_pairs_per_s": kern_pairs_per_s, "tensor_stats": tstats}
        if clocks.get("sm_mhz"):
            roofline["frac_at_sampled_clock"] = roofline["frac"] * sm_max_mhz / clocks["sm_mhz"]
    else:
        # reference-parallel kernel (k > 128, or very few queries)
        hbm = float(pk.get("hbm_gbs", 6650.0))
        if m < 16:
            bytes_per_launch = 4.0 * (k + 1) * (r1 - r0) + 4.0 * k * m + 8.0 * m
            achieved = bytes_per_launch / (kern_ms_avg * 1e-3) / 1e9
            roofline = {"bound": "hbm", "achieved": achieved, "peak": hbm, "unit": "GB/s", "frac": achieved / hbm,
                        "traffic": None, "peak_source": f"{pk_src} hbm_gbs", "kernel_ms": kern_ms_avg}
        else:
            achieved = kern_pairs_per_s * 4.0 * k / 1e12
            roofline = {"bound": "fp32", "achieved": achieved, "peak": fp32_peak_tflops, "unit": "TFLOP/s",
                        "frac": achieved / fp32_peak_tflops, "traffic": None, "kernel": "wide (reference-parallel) kernel",
                        "executed_lane_slots_per_pair": 2 * k, "kernel_ms": kern_ms_avg, "kernel_pairs_per_s": kern_pairs_per_s,
                        "peak_source": f"derived 2*128 lanes*148 SMs*{sm_max_mhz:.0f} MHz"}

    if world == 1:
        pname = {0: "lowk", 1: "wide", 2: "tensor"}[0 if (fell_back and k <= 32) else path]
        if args.flags & (nns_b200.FLAG_EXACT_FORM | nns_b200.FLAG_V0_ROUNDING):
            pname += "_exact"
        roofline["traffic"] = ncu_traffic(name, pname)
    line = {
        "metric": "pair_dist_evals_per_s", "value": value, "unit": "pairs/s", "queries_per_s": queries_per_s,
        "n_gpus": n_gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step,
        "wall_ms_per_step": wall_ms / args.steps, "higher_is_better": True,
        "scaling": "weak" if (shard == "query" and not strong_q) else "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(name, n_gpus, f"{shard}-sharded x{world}" + (" + NCCL MIN all-reduce of packed keys" if (shard == "reference" and world > 1) else " (no data-path collective)") + (f", {m} of {m_total} queries per GPU" if strong_q else ""),
                                  "flushed (256 MiB write) between timed steps"),
        "roofline": roofline, "clocks": clocks, "gpu_launches": int(launches),
        "plan": nns_b200.plan(k, m, r1 - r0, args.flags),
    }

    # ---- end to end through the host-pointer C ABI, pinned host buffers ----
    if not args.no_e2e:
        s_pin = torch.from_numpy(s_host).pin_memory()
        r_pin = torch.from_numpy(r_host[r0:r1] if shard == "reference" else r_host).pin_memory()
        out = np.empty(m, dtype=np.int32)
        ne2e = max(2, min(args.steps, 5))
        nr = r_pin.shape[0]
        for _ in range(2):
            nns_b200.search_host(k, m, nr, s_pin.data_ptr(), r_pin.data_ptr(), out)
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        for _ in range(ne2e):
            nns_b200.search_host(k, m, nr, s_pin.data_ptr(), r_pin.data_ptr(), out)
        e2e_s = (time.perf_counter() - t0) / ne2e
        te = torch.tensor([e2e_s], dtype=torch.float64, device=device)
        if world > 1:
            dist.all_reduce(te, op=dist.ReduceOp.MAX)
        e2e_s = float(te.item())
        line["e2e"] = {"value": pairs_per_step / e2e_s, "unit": "pairs/s", "ms_per_step": e2e_s * 1e3,
                       "h2d_bytes_per_step": int(s_pin.numel() * 4 + r_pin.numel() * 4), "d2h_bytes_per_step": int(m * 4),
                       "steps": ne2e, "api": "nns_b200_search_host (pinned host buffers; H2D + index build + search + D2H)"}
        assert np.array_equal(out, idx.cpu().numpy()) or shard == "reference", "host-ABI result differs from the device path"

    # ---- CPU baseline: the reference's V0 under OpenMP on a bounded sample (rank 0, N=1 only) ----
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        ms = min(m, args.cpu_sample)
        rate, info, v_idx = cpu_reference_rate(k, n, s_host[:ms], r_host, 1)
        line["cpu_baseline"] = {"value": rate, "unit": "pairs/s", "cores": info["cores"], "kind": info["kind"],
                                "sample": f"first {ms} of {m} queries x all {n} references, {info['seconds']:.2f} s",
                                "index_agreement_with_gpu": float((v_idx == idx[:ms].cpu().numpy()).mean())}

    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
