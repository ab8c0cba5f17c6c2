#!/usr/bin/env python
"""tools/overflow_cost.py -- what the tcgen05 screen's overflow hand-over costs on data it cannot
resolve (BASELINE config C5's clustered, grid-snapped, duplicated points): a single-wave query set
(148 strips of 256) against n references, planner default vs the FP32 screened kernel alone."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "nns-cuda_b200"))
import torch
import nns_b200
from nns_b200 import datagen

m, n = 148 * 256, int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 24
s, r = datagen.clustered_workload(m, n, 3, 1000)
dev = torch.device("cuda", 0)
d_q = torch.from_numpy(s).to(dev)
index = nns_b200.DeviceIndex(torch.from_numpy(r).to(dev))
st = torch.cuda.current_stream()
out = {}
for name, flags in (("planner default", 0), ("FP32 screened kernel", nns_b200.FLAG_FORCE_LOWK)):
    ms = []
    for i in range(4):
        keys = index.new_keys(m)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(st)
        index.search_keys(d_q, keys, flags, st)
        e1.record(st)
        torch.cuda.synchronize()
        ms.append(e0.elapsed_time(e1))
    out[name] = keys.clone()
    extra = nns_b200.tensor_stats() if flags == 0 else ""
    print(f"k=3 m={m} n={n} clustered: {name:22s} {min(ms[1:]):9.3f} ms  plan path {nns_b200.plan(3, m, n, flags)['path']} {extra}")
assert torch.equal(out["planner default"], out["FP32 screened kernel"])
print("keys identical")
