#!/usr/bin/env python
"""tools/topk_probe.py -- K-nearest-neighbour search through the tcgen05 screen against the FP32 list
kernel alone, device-resident inputs, at the C2 and (reduced) C3 shapes.  Prints one line per case and
checks that both paths return the same packed keys."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "nns-cuda_b200"))
import torch
import nns_b200

dev = torch.device("cuda", 0)
st = torch.cuda.current_stream()
cases = [(3, 65536, 4194304, 8), (3, 65536, 4194304, 32), (16, 65536, 4194304, 8), (16, 65536, 4194304, 32),
         (3, 8192, 1048576, 8), (64, 32768, 1048576, 8)]
if len(sys.argv) > 1:
    cases = [tuple(int(x) for x in a.split(",")) for a in sys.argv[1:]]
for k, m, n, K in cases:
    rng = np.random.default_rng(1000 + k)
    d_q = torch.from_numpy(rng.random((m, k), dtype=np.float32)).to(dev)
    index = nns_b200.DeviceIndex(torch.from_numpy(rng.random((n, k), dtype=np.float32)).to(dev))
    out = {}
    line = f"k={k:3d} m={m} n={n} K={K:2d}:"
    for name, flags in (("tensor", nns_b200.FLAG_FORCE_TENSOR), ("fp32", nns_b200.FLAG_FORCE_LOWK), ("planner", 0)):
        ms = []
        for i in range(4):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(st)
            keys = index.topk_keys(d_q, K, None, flags, st)
            e1.record(st)
            torch.cuda.synchronize()
            ms.append(e0.elapsed_time(e1))
        out[name] = keys.clone()
        line += f"  {name} {min(ms[1:]):9.3f} ms"
        if name == "tensor":
            line += f" {nns_b200.tensor_stats()}"
    same = torch.equal(out["tensor"], out["fp32"]) and torch.equal(out["planner"], out["fp32"])
    print(line, " keys identical" if same else "  KEYS DIFFER", flush=True)
    del index
