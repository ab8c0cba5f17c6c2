"""Generates tests/golden/v0_golden.npz by running the REFERENCE's own V0 (namespace v0 of
/root/reference/core.cu:11-54, compiled by oracle/build_ref.sh into oracle/_ref/libv0_ref.so) on
seeded inputs.  Needs /root/reference, so it only runs in the build container; the fixture it
writes is committed and is what pins the oracle restatement (and the CUDA path) elsewhere.

    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "nns-cuda_b200"))
sys.path.insert(0, os.path.join(ROOT, "tests"))

from conftest import make_case  # noqa: E402
from oracle import oracle  # noqa: E402

# (kind, k, m, n, seed) -- small enough to commit, wide enough to cover every k family,
# ragged sizes around the 128-point reference block, ties/duplicates and clustered data
CASES = [
    ("uniform", 3, 64, 1000, 1), ("uniform", 3, 33, 129, 2), ("uniform", 1, 40, 127, 3),
    ("uniform", 2, 17, 128, 4), ("uniform", 4, 50, 513, 5), ("uniform", 5, 31, 777, 6),
    ("uniform", 8, 20, 1025, 7), ("uniform", 16, 64, 2048, 8), ("uniform", 17, 9, 300, 9),
    ("uniform", 31, 12, 257, 10), ("uniform", 32, 16, 640, 11), ("uniform", 33, 5, 200, 12),
    ("uniform", 64, 8, 500, 13), ("uniform", 128, 16, 1000, 14), ("uniform", 3, 1, 4096, 15),
    ("grid", 3, 128, 2000, 16), ("grid", 2, 64, 999, 17), ("grid", 16, 32, 1500, 18),
    ("clustered", 3, 200, 3000, 19), ("uniform", 3, 1024, 65536, 1000),
]


def main():
    oracle.build()
    assert oracle.ref() is not None, "oracle/_ref/libv0_ref.so missing: run oracle/build_ref.sh"
    out = {}
    for ci, (kind, k, m, n, seed) in enumerate(CASES):
        s, r = make_case(kind, k, m, n, seed)
        idx = oracle.ref_v0(k, m, n, s, r)
        out[f"case{ci}_meta"] = np.array([k, m, n, seed], dtype=np.int64)
        out[f"case{ci}_kind"] = np.array(kind)
        out[f"case{ci}_idx"] = idx.astype(np.int32)
        # input checksums pin the generators too
        out[f"case{ci}_sum"] = np.array([s.astype(np.float64).sum(), r.astype(np.float64).sum()])
    # the reference's literal generator (main.cu:24-35, srand(1000)) for its first 4 shapes
    from nns_b200 import datagen

    for si, (k, m, n) in enumerate([(3, 1, 1024), (16, 1, 1024)]):
        s, r = datagen.reference_rand_sample(k, m, n, 1000)
        out[f"rand{si}_meta"] = np.array([k, m, n], dtype=np.int64)
        out[f"rand{si}_idx"] = oracle.ref_v0(k, m, n, s, r).astype(np.int32)
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "v0_golden.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
