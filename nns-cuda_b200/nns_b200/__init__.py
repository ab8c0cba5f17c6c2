"""nns_b200 -- thin ctypes binding of libnns_b200.so (include/nns_b200.h).

This module is plumbing for tests and bench.py: it mirrors the reference's callback surface
(``cudaCall(k, m, n, s_points, r_points) -> int[m]``, reference core.cu:23-29 / main.cu:74)
and exposes the device-resident building blocks on torch tensors (torch is used only for
device memory and streams).  There is NO CPU fallback: if the CUDA library is missing the
import fails, and every call fails loudly when no sm_100 GPU is usable.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import POINTER, c_char_p, c_float, c_int, c_size_t, c_uint, c_uint64, c_ulonglong, c_void_p

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
# NNS_B200_LIB: another build of the same library (tools/tensor_tune.sh's kernel variants)
LIB_PATH = os.environ.get("NNS_B200_LIB") or os.path.join(os.path.dirname(_HERE), "lib", "libnns_b200.so")

OK, ERR_INVALID, ERR_CUDA, ERR_UNSUPPORTED, ERR_NOMEM = 0, 1, 2, 3, 4
REF_BLOCK = 128
KEY_INIT = 0x7F80000000000000
FLAG_V0_ROUNDING, FLAG_FORCE_LOWK, FLAG_FORCE_WIDE, FLAG_FORCE_TENSOR, FLAG_EXACT_FORM = 1, 2, 4, 8, 16
INDEX_HEADER_FLOATS = 32
FLAG_TEST_TINY_CANDIDATES = 1 << 28


def flag_overrides(q: int = 0, warps: int = 0, stages: int = 0) -> int:
    """Tuning overrides packed into the flags word (bits 8-15 q, 16-23 warps, 24-27 stages)."""
    return (q & 0xFF) << 8 | (warps & 0xFF) << 16 | (stages & 0xF) << 24


def nns_plan_q(k: int):
    """The two register blockings (queries per thread) compiled for dimension k (csrc/nns_plan.h)."""
    return (4, 8) if k <= 4 else (4, 2) if k <= 8 else (2, 4) if k <= 16 else (2, 1)


class NnsError(RuntimeError):
    def __init__(self, status: int, msg: str):
        super().__init__(f"nns_b200 status {status}: {msg}")
        self.status = status


if not os.path.exists(LIB_PATH):
    raise ImportError(
        f"{LIB_PATH} is missing: build it with `make -C nns-cuda_b200` (or __graft_entry__.build()); "
        "there is no CPU fallback for the search path"
    )

lib = ctypes.CDLL(LIB_PATH)
_libc = ctypes.CDLL(None)
_libc.free.argtypes = [c_void_p]
_libc.free.restype = None

_fp = POINTER(c_float)
lib.nns_b200_version.restype = c_int
lib.nns_b200_last_error.restype = c_char_p
lib.nns_b200_launch_count.restype = c_ulonglong
lib.nns_b200_init.argtypes = [c_int]
lib.nns_b200_shutdown.argtypes = []
lib.nns_b200_cudaCall.argtypes = [c_int, c_int, c_int, _fp, _fp, POINTER(POINTER(c_int))]
lib.nns_b200_cudaCall.restype = None
lib.nns_b200_tensor_bound.argtypes = [c_int, c_int, ctypes.c_float, ctypes.c_float, ctypes.c_float, POINTER(ctypes.c_float)]
lib.nns_b200_search_host.argtypes = [c_int, c_int, c_int, c_void_p, c_void_p, c_void_p]
lib.nns_b200_search_host_dist.argtypes = [c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p]
lib.nns_b200_search_multi.argtypes = [c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_int, c_int]
lib.nns_b200_index_floats.argtypes = [c_int, c_int]
lib.nns_b200_index_floats.restype = c_size_t
lib.nns_b200_index_build.argtypes = [c_int, c_int, c_void_p, c_void_p, c_void_p]
lib.nns_b200_keys_init.argtypes = [c_void_p, c_int, c_void_p]
lib.nns_b200_search_keys.argtypes = [c_int, c_int, c_int, c_void_p, c_void_p, c_int, c_void_p, c_uint, c_void_p]
lib.nns_b200_keys_unpack.argtypes = [c_void_p, c_int, c_void_p, c_void_p, c_void_p]
lib.nns_b200_workspace_bytes.argtypes = [c_int, c_int, c_int]
lib.nns_b200_workspace_bytes.restype = c_size_t
lib.nns_b200_search_device.argtypes = [c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_size_t, c_uint, c_void_p]
lib.nns_b200_tensor_stats.argtypes = [POINTER(c_uint)]
lib.nns_b200_device_sms.argtypes = [c_int]
lib.nns_b200_topk_keys.argtypes = [c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_int, c_void_p, c_uint, c_void_p]
lib.nns_b200_topk_unpack.argtypes = [c_void_p, c_int, c_int, c_void_p, c_void_p, c_void_p]
lib.nns_b200_search_topk_host.argtypes = [c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p]
lib.nns_b200_index_create.argtypes = [c_int, c_int, c_void_p, c_int, POINTER(c_void_p)]
lib.nns_b200_index_search.argtypes = [c_void_p, c_int, c_void_p, c_void_p, c_void_p]
lib.nns_b200_index_size.argtypes = [c_void_p, POINTER(c_int), POINTER(c_int)]
lib.nns_b200_index_destroy.argtypes = [c_void_p]
lib.nns_b200_tree_create.argtypes = [c_int, c_int, c_void_p, c_int, POINTER(c_void_p)]
lib.nns_b200_tree_search.argtypes = [c_void_p, c_int, c_void_p, c_void_p, c_void_p]
lib.nns_b200_tree_destroy.argtypes = [c_void_p]
lib.nns_b200_sample_centre.argtypes = [c_int, c_int, c_void_p, c_void_p]
lib.nns_b200_index_build_part.argtypes = [c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_int, c_void_p]
lib.nns_b200_index_part_ranges.argtypes = [c_int, c_int, c_int, c_int, c_int, POINTER(c_size_t)]
lib.nns_b200_index_finish.argtypes = [c_int, c_int, c_void_p, c_int, c_void_p]
lib.nns_b200_plan.argtypes = [c_int, c_int, c_int, c_uint, c_int, POINTER(c_int)]


def _check(status: int) -> None:
    if status != OK:
        raise NnsError(status, (lib.nns_b200_last_error() or b"").decode(errors="replace"))


def _host_f32(a, rows: int, k: int) -> np.ndarray:
    a = np.ascontiguousarray(a, dtype=np.float32)
    if a.size != rows * k:
        raise ValueError(f"expected {rows}x{k} floats, got {a.size}")
    return a


# ---- the reference's callback surface ---------------------------------------------------------
def cudaCall(k: int, m: int, n: int, s_points, r_points) -> np.ndarray:
    """Mirror of vN::cudaCall (core.cu:23-29): returns the malloc'd int[m] as a numpy array
    (copied, then freed with free() as the reference's caller is expected to)."""
    s = _host_f32(s_points, m, k)
    r = _host_f32(r_points, n, k)
    res = POINTER(c_int)()
    lib.nns_b200_cudaCall(k, m, n, s.ctypes.data_as(_fp), r.ctypes.data_as(_fp), ctypes.byref(res))
    out = np.ctypeslib.as_array(res, shape=(max(m, 1),))[:m].copy()
    _libc.free(ctypes.cast(res, c_void_p))
    return out.astype(np.int32, copy=False)


def search_host(k: int, m: int, n: int, s_points, r_points, out: np.ndarray | None = None) -> np.ndarray:
    """nns_b200_search_host: status-checked variant; accepts numpy arrays or raw host pointers
    (ints) for s_points / r_points so pinned torch buffers can be passed without a copy."""
    sp = s_points if isinstance(s_points, int) else _host_f32(s_points, m, k).ctypes.data
    rp = r_points if isinstance(r_points, int) else _host_f32(r_points, n, k).ctypes.data
    if out is None:
        out = np.empty(m, dtype=np.int32)
    keep = (s_points, r_points)  # keep temporaries alive across the call
    _check(lib.nns_b200_search_host(k, m, n, sp, rp, out.ctypes.data))
    del keep
    return out


def search_host_dist(k: int, m: int, n: int, s_points, r_points):
    """nns_b200_search_host_dist: (indices, FP32 squared distances) for host arrays."""
    s = _host_f32(s_points, m, k)
    r = _host_f32(r_points, n, k)
    idx = np.empty(m, dtype=np.int32)
    dist = np.empty(m, dtype=np.float32)
    _check(lib.nns_b200_search_host_dist(k, m, n, s.ctypes.data, r.ctypes.data, idx.ctypes.data, dist.ctypes.data))
    return idx, dist


def search_multi(k: int, m: int, n: int, s_points, r_points, num_gpus: int = 0, shard_mode: int = 0) -> np.ndarray:
    s = _host_f32(s_points, m, k)
    r = _host_f32(r_points, n, k)
    out = np.empty(m, dtype=np.int32)
    _check(lib.nns_b200_search_multi(k, m, n, s.ctypes.data, r.ctypes.data, out.ctypes.data, num_gpus, shard_mode))
    return out


class HostIndex:
    """nns_b200_index_*: the reference set ingested once from a host array and kept resident in HBM;
    searches take host query arrays (numpy, or raw host pointers as ints) like the drop-in symbol."""

    def __init__(self, k: int, n: int, r_points, device: int = -1):
        self.k, self.n = int(k), int(n)
        rp = r_points if isinstance(r_points, int) else _host_f32(r_points, n, k).ctypes.data
        h = c_void_p()
        _check(lib.nns_b200_index_create(k, n, rp, device, ctypes.byref(h)))
        self._h = h

    def search(self, m: int, s_points, out: np.ndarray | None = None, return_dist: bool = False):
        keep = s_points
        sp = s_points if isinstance(s_points, int) else _host_f32(s_points, m, self.k).ctypes.data
        if out is None:
            out = np.empty(m, dtype=np.int32)
        dist = np.empty(m, dtype=np.float32) if return_dist else None
        _check(lib.nns_b200_index_search(self._h, m, sp, out.ctypes.data, dist.ctypes.data if return_dist else None))
        del keep
        return (out, dist) if return_dist else out

    def close(self):
        if self._h:
            _check(lib.nns_b200_index_destroy(self._h))
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def search_topk_host(k: int, m: int, n: int, K: int, s_points, r_points, return_dist: bool = True):
    """nns_b200_search_topk_host: (int32[m][K] indices, float32[m][K] squared distances), ascending."""
    s = _host_f32(s_points, m, k)
    r = _host_f32(r_points, n, k)
    idx = np.empty((m, K), dtype=np.int32)
    dist = np.empty((m, K), dtype=np.float32) if return_dist else None
    _check(lib.nns_b200_search_topk_host(k, m, n, K, s.ctypes.data, r.ctypes.data, idx.ctypes.data,
                                         dist.ctypes.data if return_dist else None))
    return (idx, dist) if return_dist else idx


class HostTree:
    """nns_b200_tree_*: exact nearest-neighbour search through a bucketed KD-tree (k <= 32) built from a host
    array; searches take host query arrays and return V0's indices."""

    def __init__(self, k: int, n: int, r_points, device: int = -1):
        self.k, self.n = int(k), int(n)
        r = _host_f32(r_points, n, k)
        h = c_void_p()
        _check(lib.nns_b200_tree_create(k, n, r.ctypes.data, device, ctypes.byref(h)))
        self._h = h

    def search(self, m: int, s_points, return_dist: bool = False):
        s = _host_f32(s_points, m, self.k)
        out = np.empty(m, dtype=np.int32)
        dist = np.empty(m, dtype=np.float32) if return_dist else None
        _check(lib.nns_b200_tree_search(self._h, m, s.ctypes.data, out.ctypes.data, dist.ctypes.data if return_dist else None))
        return (out, dist) if return_dist else out

    def close(self):
        if self._h:
            _check(lib.nns_b200_tree_destroy(self._h))
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def sample_centre(k: int, n: int, r_points) -> np.ndarray:
    """nns_b200_sample_centre: the common centre of the tcgen05 operand images when several GPUs build
    slices of one index (host arithmetic over <= 4096 sample rows)."""
    rp = r_points if isinstance(r_points, int) else _host_f32(r_points, n, k).ctypes.data
    out = np.zeros(max(k, 1), dtype=np.float32)
    _check(lib.nns_b200_sample_centre(k, n, rp, out.ctypes.data))
    return out


def device_sms(device: int = -1) -> int:
    return int(lib.nns_b200_device_sms(device))


def plan(k: int, m: int, n: int, flags: int = 0, num_sms: int = 148) -> dict:
    p = (c_int * 8)()
    _check(lib.nns_b200_plan(k, m, n, flags, num_sms, p))
    names = ("path", "q", "warps", "stages", "query_blocks", "splits", "blocks_per_split", "smem")
    return dict(zip(names, list(p)))


def tensor_kp(k: int) -> int:
    """Contraction length (BF16 columns) the tcgen05 screen uses for dimension k: mirror of
    tensor_geom() in csrc/tensor_search.cu (3k split-precision columns up to k = 42, + 3 norm columns)."""
    ndata = 3 * k if k <= 42 else k
    if ndata + 3 <= 16:
        return 16
    if ndata + 3 <= 32:
        return 32
    if ndata + 3 <= 64:
        return 64
    if ndata <= 64:
        return 80
    if ndata + 3 <= 128:
        return 128
    if ndata <= 128:
        return 144
    return 64 * ((ndata + 3 + 63) // 64)


def tensor_bound(k: int, mode: int, a: float, rmax: float, rmax_sampled: float) -> dict:
    """The tcgen05 screen's error bound as the library evaluates it (host-side pure function; mode 0 = default BF16 layout,
    2 = F16 operands and accumulators): {"E", "s" reference scale, "t" query scale (0: cannot be screened), "kp"}."""
    out = (ctypes.c_float * 4)()
    _check(lib.nns_b200_tensor_bound(int(k), int(mode), ctypes.c_float(a), ctypes.c_float(rmax), ctypes.c_float(rmax_sampled), out))
    return {"E": float(out[0]), "s": float(out[1]), "t": float(out[2]), "kp": int(out[3])}


def tensor_stats() -> dict:
    out = (c_uint * 4)()
    _check(lib.nns_b200_tensor_stats(out))
    # [3]: contraction columns of the operand images the index holds | precision mode << 16 (2 = F16 operands + accumulators)
    return {"candidates": int(out[0]), "overflow": int(out[1]), "capacity": int(out[2]), "kp": int(out[3]) & 0xFFFF,
            "mode": "f16" if (int(out[3]) >> 16) == 2 else "bf16"}


def index_floats(k: int, n: int) -> int:
    return int(lib.nns_b200_index_floats(k, n))


def launch_count() -> int:
    return int(lib.nns_b200_launch_count())


def init(device: int = -1) -> None:
    _check(lib.nns_b200_init(device))


def shutdown() -> None:
    _check(lib.nns_b200_shutdown())


# ---- device-resident API on torch tensors (torch = memory + streams only) ----------------------
def _stream_ptr(stream=None):
    import torch

    st = stream if stream is not None else torch.cuda.current_stream()
    return c_void_p(st.cuda_stream)


class DeviceIndex:
    """Build once, query many: the tiled-SoA reference index resident in HBM
    (replaces the per-call upload + transpose of core.cu:364-370)."""

    def __init__(self, refs, k: int | None = None, index_base: int = 0, stream=None):
        import torch

        assert refs.is_cuda and refs.dtype == torch.float32 and refs.is_contiguous()
        self.n = refs.shape[0]
        self.k = int(k if k is not None else refs.shape[1])
        self.index_base = int(index_base)
        self.device = refs.device
        nfl = index_floats(self.k, self.n)
        self.index = torch.empty(max(nfl, 4), dtype=torch.float32, device=refs.device)
        with torch.cuda.device(self.device):
            _check(lib.nns_b200_index_build(self.k, self.n, refs.data_ptr(), self.index.data_ptr(), _stream_ptr(stream)))

    @classmethod
    def from_built(cls, index, k: int, n: int, index_base: int = 0):
        """Wrap an index tensor that was built elsewhere (nns_b200_index_build_part + exchange + finish)."""
        self = cls.__new__(cls)
        self.n, self.k, self.index_base, self.device, self.index = int(n), int(k), int(index_base), index.device, index
        return self

    def new_keys(self, m: int, stream=None):
        import torch

        keys = torch.empty(max(m, 1), dtype=torch.int64, device=self.device)
        with torch.cuda.device(self.device):
            _check(lib.nns_b200_keys_init(keys.data_ptr(), m, _stream_ptr(stream)))
        return keys

    def search_keys(self, queries, keys, flags: int = 0, stream=None):
        """keys[i] = min(keys[i], key(best dist, index_base + j)) -- the hot path."""
        import torch

        assert queries.is_cuda and queries.dtype == torch.float32 and queries.is_contiguous()
        m = queries.shape[0]
        with torch.cuda.device(self.device):
            _check(lib.nns_b200_search_keys(self.k, m, self.n, queries.data_ptr(), self.index.data_ptr(),
                                            self.index_base, keys.data_ptr(), flags, _stream_ptr(stream)))
        return keys

    def topk_keys(self, queries, K: int, keys=None, flags: int = 0, stream=None):
        """uint64-as-int64 [m][K] ascending packed keys of the K nearest references (accumulates into
        `keys` when given: shards / GPUs merge like the 1-NN keys)."""
        import torch

        assert queries.is_cuda and queries.dtype == torch.float32 and queries.is_contiguous()
        m = queries.shape[0]
        if keys is None:
            keys = torch.empty((max(m, 1), K), dtype=torch.int64, device=self.device)
            with torch.cuda.device(self.device):
                _check(lib.nns_b200_keys_init(keys.data_ptr(), m * K, _stream_ptr(stream)))
        with torch.cuda.device(self.device):
            _check(lib.nns_b200_topk_keys(self.k, m, self.n, K, queries.data_ptr(), self.index.data_ptr(), self.index_base,
                                          keys.data_ptr(), flags, _stream_ptr(stream)))
        return keys

    def topk(self, queries, K: int, flags: int = 0, stream=None):
        import torch

        m = queries.shape[0]
        keys = self.topk_keys(queries, K, None, flags, stream)
        idx = torch.empty((max(m, 1), K), dtype=torch.int32, device=self.device)
        dist = torch.empty((max(m, 1), K), dtype=torch.float32, device=self.device)
        with torch.cuda.device(self.device):
            _check(lib.nns_b200_topk_unpack(keys.data_ptr(), m, K, idx.data_ptr(), dist.data_ptr(), _stream_ptr(stream)))
        return idx[:m], dist[:m]

    def search(self, queries, flags: int = 0, stream=None, return_dist: bool = False):
        import torch

        m = queries.shape[0]
        keys = self.new_keys(m, stream)
        self.search_keys(queries, keys, flags, stream)
        return unpack_keys(keys, m, stream, return_dist)


def unpack_keys(keys, m: int, stream=None, return_dist: bool = False):
    import torch

    idx = torch.empty(max(m, 1), dtype=torch.int32, device=keys.device)
    dist = torch.empty(max(m, 1), dtype=torch.float32, device=keys.device) if return_dist else None
    with torch.cuda.device(keys.device):
        _check(lib.nns_b200_keys_unpack(keys.data_ptr(), m, idx.data_ptr(),
                                        dist.data_ptr() if dist is not None else None, _stream_ptr(stream)))
    return (idx[:m], dist[:m]) if return_dist else idx[:m]


def search_device(queries, refs, flags: int = 0, stream=None):
    """One-shot device search: index_build + search + unpack (nns_b200_search_device)."""
    import torch

    m, k = queries.shape
    n = refs.shape[0]
    ws = torch.empty(int(lib.nns_b200_workspace_bytes(k, m, n)), dtype=torch.uint8, device=queries.device)
    idx = torch.empty(max(m, 1), dtype=torch.int32, device=queries.device)
    with torch.cuda.device(queries.device):
        _check(lib.nns_b200_search_device(k, m, n, queries.data_ptr(), refs.data_ptr(), idx.data_ptr(),
                                          ws.data_ptr(), ws.numel(), flags, _stream_ptr(stream)))
    return idx[:m]
