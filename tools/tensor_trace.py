#!/usr/bin/env python
"""tools/tensor_trace.py LIB [workload] -- print the clock64() timeline of CTA (0,0) of the tcgen05
screen (a libnns_b200 built with -DNNS_T_TRACE by tools/tensor_tune.sh): per tile, when the producer
got its free stage, when the MMA issuer got the drained accumulator / the landed stage / had issued,
and when epilogue warp 0 released / received an accumulator."""
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
print("tile  M:issued | E:acc ready (the tile's team of 8 warps, relative to M:issued of the tile) | E:released (same)")
for i in range(32):
    mi = int(out[3][i])
    team = i % 2 if out[4 + 8][1] != 0 or out[4 + 8][0] != 0 else 0
    print(f"{i:4d} {mi - t0:8d} | " + " ".join(f"{int(out[4 + team * 8 + e][i] - mi):5d}" for e in range(8))
          + " | " + " ".join(f"{int(out[20 + team * 8 + e][i] - mi):5d}" for e in range(8)))
d = np.diff(out[3])
print("MMA issue period per tile: median %.0f clk" % np.median(d))
