"""Multi-GPU partitioning of the brute-force search (one process per GPU, torch.distributed).

The path shards both ways:
  * query-sharded     -- rank g answers queries [q0, q1) against the full reference set; no
                         data-path collective (optionally an all-gather of the int32 slices);
  * reference-sharded -- rank g owns a contiguous slice of whole 128-point reference blocks
                         (the reference's contiguous ceil(n/G) slices, core.cu:781-791, without the
                         <= 0 tail defect D9) and searches all queries with index_base = r0; the
                         per-query packed (dist, idx) keys are merged with ONE integer MIN
                         all-reduce, which is exact and order-independent (lowest index on ties),
                         replacing the reference's host-side merge (core.cu:821-852, defects D3/D10).
Pure host logic here; the kernels are in csrc/.
"""
from __future__ import annotations

import numpy as np

REF_BLOCK = 128
KEY_INIT = 0x7F80000000000000


def query_shard(m: int, world: int, rank: int) -> tuple[int, int]:
    per = (m + world - 1) // world
    q0 = min(m, rank * per)
    return q0, min(m, q0 + per)


def reference_shard(n: int, world: int, rank: int) -> tuple[int, int]:
    blocks = (n + REF_BLOCK - 1) // REF_BLOCK
    per = ((blocks + world - 1) // world) * REF_BLOCK
    r0 = min(n, rank * per)
    return r0, min(n, r0 + per)


def pack_keys(dist_f32: np.ndarray, idx: np.ndarray) -> np.ndarray:
    """(float_bits(dist) << 32) | idx as int64 (dist >= +0, so signed order == unsigned order)."""
    bits = np.ascontiguousarray(dist_f32, dtype=np.float32).view(np.uint32).astype(np.uint64)
    return ((bits << np.uint64(32)) | np.asarray(idx).astype(np.uint32).astype(np.uint64)).astype(np.int64)


def unpack_keys(keys: np.ndarray) -> tuple[np.ndarray, np.ndarray]:
    u = np.asarray(keys).astype(np.uint64)
    return (u & np.uint64(0xFFFFFFFF)).astype(np.int32), (u >> np.uint64(32)).astype(np.uint32).view(np.float32)


def allreduce_min_keys(keys):
    """The one exchange step of the reference-sharded path: in-place MIN all-reduce of the int64
    packed keys (NCCL over NVLink on GPUs; gloo in the CPU tests)."""
    import torch.distributed as dist

    dist.all_reduce(keys, op=dist.ReduceOp.MIN)
    return keys


def allgather_indices(idx_local, m: int, world: int):
    """Optional for the query-sharded path: replicate the int32 result slices on every rank."""
    import torch
    import torch.distributed as dist

    per = (m + world - 1) // world
    pad = torch.zeros(per, dtype=idx_local.dtype, device=idx_local.device)
    pad[: idx_local.numel()] = idx_local
    out = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(out, pad)
    return torch.cat(out)[:m]
