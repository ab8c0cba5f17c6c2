"""tools/e2e_probe.py -- C3 through nns_b200_search_host, pinned and pageable, call by call (NNS_B200_TRACE=1 prints the
host-side phases; NNS_B200_INGEST_CHUNK_MB sets the ingest chunk).  Found the idle-stream cudaStreamSynchronize stall
that host_state.h::stream_drain works around."""
import os, sys, time, numpy as np
sys.path.insert(0, 'nns-cuda_b200'); sys.path.insert(0, '.')
import torch, nns_b200
from nns_b200 import datagen
k, m, n = 16, 262144, 16777216
s = datagen.uniform_points(m, k, 1000, 0); r = datagen.uniform_points(n, k, 1000, 1)
sp, rp = torch.from_numpy(s).pin_memory(), torch.from_numpy(r).pin_memory()
out = np.empty(m, np.int32)
d = torch.empty_like(rp, device='cuda')
for _ in range(2):
    torch.cuda.synchronize(); t0 = time.perf_counter(); d.copy_(rp, non_blocking=True); torch.cuda.synchronize(); t = time.perf_counter() - t0
print('H2D 1 GiB pinned: %.1f ms = %.1f GB/s' % (t * 1e3, rp.numel() * 4 / t / 1e9))
del d
ts = []
for i in range(6):
    t0 = time.perf_counter(); nns_b200.search_host(k, m, n, sp.data_ptr(), rp.data_ptr(), out); ts.append((time.perf_counter() - t0) * 1e3)
print('chunk env', os.environ.get('NNS_B200_INGEST_CHUNK_MB'), 'pinned e2e ms:', [round(x, 1) for x in ts], nns_b200.tensor_stats())
ts = []
for i in range(4):
    t0 = time.perf_counter(); nns_b200.search_host(k, m, n, s, r, out); ts.append((time.perf_counter() - t0) * 1e3)
print('pageable e2e ms:', [round(x, 1) for x in ts])
