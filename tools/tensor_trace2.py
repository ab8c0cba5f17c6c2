#!/usr/bin/env python
"""tools/tensor_trace2.py LIB [k,m,n] -- clock64() timeline of CTA (0,0) of the tcgen05 screen with the
two-units-per-trip epilogue (a libnns_b200 built with -DNNS_T_TRACE=<first unit> by tools/tensor_tune.sh):
per accumulator unit, when the MMA issuer committed it and -- for warp 0 of the unit's epilogue team --
when the warp started waiting for it, got it, and released it."""
import ctypes, os, sys
import numpy as np
os.environ["NNS_B200_LIB"] = os.path.abspath(sys.argv[1])
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "nns-cuda_b200"))
import torch
import nns_b200
from nns_b200 import datagen
k, m, n = (3, 148 * 256, 1 << 20) if len(sys.argv) < 3 else tuple(int(x) for x in sys.argv[2].split(","))
s = datagen.uniform_points(m, k, 1000, 0)
r = datagen.uniform_points(n, k, 1000, 1)
dev = torch.device("cuda", 0)
d_q = torch.from_numpy(s).to(dev)
index = nns_b200.DeviceIndex(torch.from_numpy(r).to(dev))
keys = index.new_keys(m)
st = torch.cuda.current_stream()
for _ in range(3):
    nns_b200._check(nns_b200.lib.nns_b200_keys_init(keys.data_ptr(), m, st.cuda_stream))
    index.search_keys(d_q, keys, nns_b200.FLAG_FORCE_TENSOR, st)
torch.cuda.synchronize()
out = np.zeros((36, 32), dtype=np.int64)
rc = nns_b200.lib.nns_b200_debug_trace(ctypes.c_void_p(out.ctypes.data))
assert rc == 0, rc
t0 = out[3].min()
print("unit  issued(abs) | per warp of the unit's team, relative to `issued`: wait-start / ready / released   (warps 0..7)")
for i in range(32):
    mi = int(out[3][i])
    ws = " ".join(f"{int(out[24 + e][i] - mi):5d}" for e in range(0, 8, 2))
    rd = " ".join(f"{int(out[4 + e][i] - mi):5d}" for e in range(0, 8, 2))
    rl = " ".join(f"{int(out[12 + e][i] - mi):5d}" for e in range(0, 8, 2))
    iss = f"{int(out[1][i] - mi):5d} {int(out[2][i] - mi):5d}"  # the issuer: started waiting for the buffer / got it (and the B stage)
    print(f"{i:4d} {mi - t0:8d} | issuer wait-start, go {iss} | wait {ws} | ready {rd} | released {rl}")
d = np.diff(out[3])
print("MMA issue period per unit: median %.0f clk, mean %.0f" % (np.median(d), d.mean()))
