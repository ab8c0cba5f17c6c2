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


def gather_built_index(k: int, n: int, r_host, rank: int, world: int, device, stream=None):
    """Query-sharded search needs the whole reference index on every GPU.  Instead of `world` uploads
    of the full set (what the reference does, core.cu:793-818), rank g uploads only slice g over its
    own PCIe link, builds that slice of the index -- FP32 blocks and tcgen05 operand images, with a
    common centre -- in place (nns_b200_index_build_part), and the built slices are exchanged with
    in-place NCCL all-gathers over NVLink; nns_b200_index_finish folds the per-slice maxima.

    r_host: the host reference array [n][k] (numpy, or a pinned torch CPU tensor).  Returns a
    DeviceIndex over n_pad >= n references (equal-sized slices; the padding is NaN and never wins).
    """
    import ctypes

    import torch
    import torch.distributed as dist

    import nns_b200
    from nns_b200 import lib, _check

    r_t = r_host if isinstance(r_host, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(r_host, dtype=np.float32))
    r_np = r_t.numpy()
    blocks = (n + REF_BLOCK - 1) // REF_BLOCK
    per_blocks = (blocks + world - 1) // world
    n_pad = per_blocks * world * REF_BLOCK
    j0 = rank * per_blocks * REF_BLOCK
    cn = max(0, min(n - j0, per_blocks * REF_BLOCK))
    st = stream if stream is not None else torch.cuda.current_stream()
    index = torch.empty(nns_b200.index_floats(k, n_pad), dtype=torch.float32, device=device)
    with torch.cuda.device(device), torch.cuda.stream(st):
        d_part = r_t[j0:j0 + cn].to(device, non_blocking=True) if cn > 0 else torch.empty((0, k), dtype=torch.float32, device=device)
        centre = nns_b200.sample_centre(k, n, r_np) if k <= 509 else None
        _check(lib.nns_b200_index_build_part(k, n_pad, j0, cn, per_blocks, d_part.data_ptr(), index.data_ptr(),
                                             centre.ctypes.data if centre is not None else None, rank, ctypes.c_void_p(st.cuda_stream)))
        if world > 1:
            rg = (ctypes.c_size_t * 6)()
            _check(lib.nns_b200_index_part_ranges(k, n_pad, 0, per_blocks, 0, rg))
            b_off, b_bytes, i_off, i_bytes, h_off, s_off = (int(x) for x in rg)
            raw = index.view(torch.uint8)

            def gather(off, nbytes):
                whole = raw[off:off + nbytes * world]
                dist.all_gather_into_tensor(whole, whole[rank * nbytes:(rank + 1) * nbytes])

            gather(b_off, b_bytes)
            gather(h_off, 4)
            if i_bytes:
                gather(i_off, i_bytes)
                gather(s_off, 4)        # partial max |r'|^2
                gather(s_off + 64, 4)   # partial flags
        _check(lib.nns_b200_index_finish(k, n_pad, index.data_ptr(), world, ctypes.c_void_p(st.cuda_stream)))
    return nns_b200.DeviceIndex.from_built(index, k, n_pad)
