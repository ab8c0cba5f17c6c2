"""CPU tests (-m "not gpu") of the N > 1 host logic with torch.distributed/gloo, world_size 2:
shard geometry, packed-key MIN all-reduce (reference-sharded) and all-gather (query-sharded).
The per-shard partial results are produced by the V0 oracle here (test infrastructure); on GPUs
the same keys come out of nns_b200_search_keys."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import make_case


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _v0_keys(oracle, k, s, r, base):
    """packed keys of V0's answer over one reference shard (dist recomputed in FP32 like V0)."""
    from nns_b200 import sharding

    m, n = s.shape[0], r.shape[0]
    if n == 0:
        return np.full(m, sharding.KEY_INIT, dtype=np.uint64).astype(np.int64)
    idx = oracle.v0(k, m, n, s, r)
    d = np.zeros(m, dtype=np.float32)
    for t in range(k):
        diff = s[:, t] - r[idx, t]
        d = d + diff * diff
    keys = sharding.pack_keys(d, idx + base)
    keys[~(d < np.inf)] = np.int64(sharding.KEY_INIT)  # V0 never selects a non-finite distance
    return keys


def _worker(rank, world, port, kind, k, m, n, seed, out_dir):
    import sys

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, root)
    sys.path.insert(0, os.path.join(root, "nns-cuda_b200"))
    from nns_b200 import sharding
    from oracle import oracle

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    s, r = make_case(kind, k, m, n, seed)
    # reference-sharded: one MIN all-reduce of packed keys
    r0, r1 = sharding.reference_shard(n, world, rank)
    keys = torch.from_numpy(_v0_keys(oracle, k, s, r[r0:r1], r0))
    sharding.allreduce_min_keys(keys)
    idx_ref, _ = sharding.unpack_keys(keys.numpy())
    # query-sharded: no exchange for the result, optional all-gather
    q0, q1 = sharding.query_shard(m, world, rank)
    local = torch.from_numpy(oracle.v0(k, q1 - q0, n, s[q0:q1], r))
    idx_q = sharding.allgather_indices(local, m, world).numpy()
    if rank == 0:
        np.savez(os.path.join(out_dir, "out.npz"), idx_ref=idx_ref, idx_q=idx_q)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("kind,k,m,n", [("grid", 3, 257, 1000), ("uniform", 16, 64, 300), ("clustered", 3, 128, 129)])
def test_world2_sharded_results_equal_v0(oracle, tmp_path, kind, k, m, n):
    port = _free_port()
    mp.spawn(_worker, args=(2, port, kind, k, m, n, 5, str(tmp_path)), nprocs=2, join=True)
    z = np.load(tmp_path / "out.npz")
    s, r = make_case(kind, k, m, n, 5)
    v = oracle.v0(k, m, n, s, r)
    assert np.array_equal(z["idx_ref"], v)  # exact lowest-index merge, independent of shard count
    assert np.array_equal(z["idx_q"], v)


def test_shard_geometry():
    from nns_b200 import sharding

    for n in (0, 1, 5, 127, 128, 129, 1000, 16777216):
        for world in (1, 2, 3, 4, 8):
            cover = []
            for rank in range(world):
                r0, r1 = sharding.reference_shard(n, world, rank)
                assert 0 <= r0 <= r1 <= n and (r0 % 128 == 0 or r0 == n)  # never a negative tail (reference defect D9)
                cover.append((r0, r1))
            assert cover[0][0] == 0 and cover[-1][1] == n
            assert all(cover[i][1] == cover[i + 1][0] for i in range(world - 1))
    for m in (0, 1, 7, 65536):
        for world in (1, 2, 8):
            qs = [sharding.query_shard(m, world, r) for r in range(world)]
            assert qs[0][0] == 0 and qs[-1][1] == m and all(qs[i][1] == qs[i + 1][0] for i in range(world - 1))


def test_key_packing_orders_like_dist_then_index():
    from nns_b200 import sharding

    d = np.array([0.0, 1e-30, 0.5, 0.5, np.inf], np.float32)
    i = np.array([9, 3, 7, 2, 0])
    keys = sharding.pack_keys(d, i)
    assert list(np.argsort(keys, kind="stable")) == [0, 1, 3, 2, 4]
    assert keys[4] == np.int64(sharding.KEY_INIT) and (keys >= 0).all()
    idx, dist_back = sharding.unpack_keys(keys)
    assert np.array_equal(idx, i) and np.array_equal(dist_back, d)
