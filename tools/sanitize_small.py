"""tools/sanitize_small.py -- one small call of every kernel family, for compute-sanitizer:
    compute-sanitizer --tool memcheck python tools/sanitize_small.py
(the sizes keep a memcheck run to a few minutes; results are checked against the V0 oracle as well)."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "nns-cuda_b200")); sys.path.insert(0, ROOT)
import torch
import nns_b200
from nns_b200 import datagen
from oracle import oracle

def check(name, k, m, n, s, r, g):
    v, _ = oracle.v0_omp(k, m, n, s, r)
    rep = oracle.check_tie_rule(k, m, n, s, r, g, v, 1e-5)
    assert rep["violations"] == 0, (name, rep)
    print("ok", name, flush=True)

dev = torch.device("cuda", 0)
for (k, m, n, flags, name) in [(3, 600, 5000, nns_b200.FLAG_FORCE_LOWK, "lowk filter"), (3, 600, 5000, nns_b200.FLAG_FORCE_LOWK | nns_b200.FLAG_EXACT_FORM, "lowk exact"),
                               (16, 300, 3000, nns_b200.FLAG_FORCE_WIDE, "wide"), (3, 700, 6000, nns_b200.FLAG_FORCE_TENSOR, "tensor k=3 split"),
                               (16, 700, 6000, nns_b200.FLAG_FORCE_TENSOR, "tensor k=16 probe"), (40, 300, 4000, nns_b200.FLAG_FORCE_TENSOR, "tensor k=40 TS"),
                               (128, 300, 4000, nns_b200.FLAG_FORCE_TENSOR, "tensor k=128"), (200, 300, 3000, nns_b200.FLAG_FORCE_TENSOR, "long-K 256 rows"),
                               (400, 200, 3000, nns_b200.FLAG_FORCE_TENSOR, "long-K 128 rows")]:
    s = datagen.uniform_points(m, k, 5, 0); r = datagen.uniform_points(n, k, 5, 1)
    idx = nns_b200.DeviceIndex(torch.from_numpy(r).to(dev)).search(torch.from_numpy(s).to(dev), flags).cpu().numpy()
    check(name, k, m, n, s, r, idx)
k, m, n = 3, 500, 70000
s, r = datagen.clustered_workload(m, n, k, 3)
check("search_host (chunks, staging)", k, m, n, s, r, nns_b200.search_host(k, m, n, s, r))
h = nns_b200.HostIndex(k, n, r); check("index handle", k, m, n, s, r, h.search(m, s)); h.close()
for mode in (0, 1):
    check(f"search_multi mode {mode}", k, m, n, s, r, nns_b200.search_multi(k, m, n, s, r, 0, mode))
ti, td = nns_b200.search_topk_host(k, m, n, 8, s, r); wi, wd = oracle.v0_topk(k, m, n, 8, s, r)
assert np.array_equal(ti, wi); print("ok topk", flush=True)
for hb in ("0", "1"):
    os.environ["NNS_B200_TREE_HOST_BUILD"] = hb
    t = nns_b200.HostTree(k, n, r); check(f"tree (host build {hb})", k, m, n, s, r, t.search(m, s)); t.close()
print("sanitize_small done")
