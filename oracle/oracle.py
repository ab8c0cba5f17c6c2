"""oracle.py -- ctypes loader for the CPU oracle.  TEST INFRASTRUCTURE ONLY: imported by tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs, never by the
product path (nns-cuda_b200/).  See v0_oracle.c for what is restated and how it is pinned."""
from __future__ import annotations

import ctypes
import os
import subprocess
from ctypes import POINTER, c_double, c_float, c_int, c_long, c_void_p

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
PORT_SO = os.path.join(HERE, "liboracle_v0.so")
REF_SO = os.path.join(HERE, "_ref", "libv0_ref.so")
REF_PROBE = os.path.join(HERE, "_ref", "ref_gpu_probe")


def build(force: bool = False) -> None:
    """Compile the restatement (always possible: gcc) and, when /root/reference is present,
    the real reference V0 into oracle/_ref/ (build_ref.sh)."""
    if force or not os.path.exists(PORT_SO) or os.path.getmtime(PORT_SO) < os.path.getmtime(os.path.join(HERE, "v0_oracle.c")):
        subprocess.check_call(["make", "-C", HERE, "liboracle_v0.so"], stdout=subprocess.DEVNULL)
    probe_stale = (not os.path.exists(REF_PROBE)
                   or os.path.getmtime(REF_PROBE) < os.path.getmtime(os.path.join(HERE, "ref_gpu_probe.cu"))
                   or os.path.getmtime(REF_SO) < os.path.getmtime(os.path.join(HERE, "ref_v0_shim_tail.inc"))) if os.path.exists(REF_SO) else True
    if os.path.exists("/root/reference/core.cu") and (force or not os.path.exists(REF_SO) or probe_stale):
        env = dict(os.environ, BUILD_REF_MAIN="0")
        subprocess.check_call([os.path.join(HERE, "build_ref.sh")], env=env, stdout=subprocess.DEVNULL)


_port = None
_ref = None


def port():
    global _port
    if _port is None:
        build()
        lib = ctypes.CDLL(PORT_SO)
        lib.oracle_v0_search.argtypes = [c_int, c_int, c_int, c_void_p, c_void_p, c_void_p]
        lib.oracle_v0_search.restype = None
        lib.oracle_v0_cudaCall.argtypes = [c_int, c_int, c_int, c_void_p, c_void_p, POINTER(POINTER(c_int))]
        lib.oracle_v0_cudaCall.restype = None
        lib.oracle_v0_search_omp.argtypes = [c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_int]
        lib.oracle_v0_search_omp.restype = c_int
        lib.oracle_num_threads.restype = c_int
        lib.oracle_libc_rand_fill.argtypes = [c_void_p, c_long]
        lib.oracle_libc_rand_fill.restype = None
        lib.oracle_v0_topk.argtypes = [c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p]
        lib.oracle_v0_topk.restype = None
        lib.oracle_set_threads.argtypes = [c_int]
        lib.oracle_set_threads.restype = None
        lib.oracle_check_tie_rule.argtypes = [c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_double, POINTER(c_long)]
        lib.oracle_check_tie_rule.restype = c_long
        _port = lib
    return _port


def ref():
    """The reference's own V0 (compiled from /root/reference/core.cu:11-54), or None."""
    global _ref
    if _ref is None and os.path.exists(REF_SO):
        lib = ctypes.CDLL(REF_SO)
        lib.ref_v0_cudaCall.argtypes = [c_int, c_int, c_int, c_void_p, c_void_p, POINTER(POINTER(c_int))]
        lib.ref_v0_cudaCall.restype = None
        lib.ref_v0_search_omp.argtypes = [c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_int]
        lib.ref_v0_search_omp.restype = c_int
        lib.ref_num_threads.restype = c_int
        if hasattr(lib, "ref_set_threads"):
            lib.ref_set_threads.argtypes = [c_int]
            lib.ref_set_threads.restype = None
        _ref = lib
    return _ref


def set_threads(n: int) -> int:
    """Team size of the OpenMP wrappers (both libraries share one libgomp).  torchrun exports
    OMP_NUM_THREADS=1, so the bench legs that time V0 on "all host cores" say so explicitly."""
    n = int(n) if n and n > 0 else (len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1))
    port().oracle_set_threads(n)
    r = ref()
    if r is not None and hasattr(r, "ref_set_threads"):
        r.ref_set_threads(n)
    return n


def _f32(a, rows, k):
    a = np.ascontiguousarray(a, dtype=np.float32)
    assert a.size == rows * k, (a.shape, rows, k)
    return a


def v0(k, m, n, s, r) -> np.ndarray:
    """Serial restatement of V0 (core.cu:31-52)."""
    s, r = _f32(s, m, k), _f32(r, n, k)
    out = np.zeros(m, dtype=np.int32)
    port().oracle_v0_search(k, m, n, s.ctypes.data, r.ctypes.data, out.ctypes.data)
    return out


def v0_omp(k, m, n, s, r, chunk: int = 8):
    """V0 per query chunk under OpenMP; returns (indices, threads)."""
    s, r = _f32(s, m, k), _f32(r, n, k)
    out = np.zeros(m, dtype=np.int32)
    threads = port().oracle_v0_search_omp(k, m, n, s.ctypes.data, r.ctypes.data, out.ctypes.data, chunk)
    return out, threads


def reference_table_inputs(shapes, seed: int = 1000):
    """Inputs of the reference's benchmark table (main.cu:54, 62-71): srand(seed) once, then for every
    (k, m, n) the queries followed by the references from the same libc stream.  Yields (k, m, n, s, r)."""
    libc = ctypes.CDLL(None)
    libc.srand(ctypes.c_uint(seed))
    for (k, m, n) in shapes:
        s = np.empty((m, k), dtype=np.float32)
        r = np.empty((n, k), dtype=np.float32)
        port().oracle_libc_rand_fill(s.ctypes.data, m * k)
        port().oracle_libc_rand_fill(r.ctypes.data, n * k)
        yield k, m, n, s, r


def v0_topk(k, m, n, K, s, r):
    """K nearest neighbours in V0's arithmetic, ordered by (distance, index): (int32[m][K], float32[m][K])."""
    s, r = _f32(s, m, k), _f32(r, n, k)
    idx = np.zeros((m, K), dtype=np.int32)
    dist = np.zeros((m, K), dtype=np.float32)
    port().oracle_v0_topk(k, m, n, K, s.ctypes.data, r.ctypes.data, idx.ctypes.data, dist.ctypes.data)
    return idx, dist


def ref_v0(k, m, n, s, r) -> np.ndarray:
    """The reference's unmodified v0::cudaCall (malloc'd result copied and freed)."""
    lib = ref()
    assert lib is not None, "oracle/_ref/libv0_ref.so not built"
    s, r = _f32(s, m, k), _f32(r, n, k)
    res = POINTER(c_int)()
    lib.ref_v0_cudaCall(k, m, n, s.ctypes.data, r.ctypes.data, ctypes.byref(res))
    out = np.ctypeslib.as_array(res, shape=(max(m, 1),))[:m].copy().astype(np.int32)
    libc = ctypes.CDLL(None)
    libc.free.argtypes = [c_void_p]
    libc.free(ctypes.cast(res, c_void_p))
    return out


def ref_v0_omp(k, m, n, s, r, chunk: int = 8):
    lib = ref()
    assert lib is not None
    s, r = _f32(s, m, k), _f32(r, n, k)
    out = np.zeros(m, dtype=np.int32)
    threads = lib.ref_v0_search_omp(k, m, n, s.ctypes.data, r.ctypes.data, out.ctypes.data, chunk)
    return out, threads


def check_tie_rule(k, m, n, s, r, engine_idx, v0_idx=None, rel_tol: float = 1e-5) -> dict:
    """North-star acceptance rule in FP64.  Returns the counters."""
    s, r = _f32(s, m, k), _f32(r, n, k)
    g = np.ascontiguousarray(engine_idx, dtype=np.int32)
    v = None if v0_idx is None else np.ascontiguousarray(v0_idx, dtype=np.int32)
    counts = (c_long * 6)()
    viol = port().oracle_check_tie_rule(k, m, n, s.ctypes.data, r.ctypes.data, g.ctypes.data,
                                        None if v is None else v.ctypes.data, rel_tol, counts)
    return {
        "violations": int(viol),
        "outside_band": int(counts[0]),
        "higher_index_on_exact_tie": int(counts[1]),
        "oracle_anomalies": int(counts[2]),
        "near_tie_accepted": int(counts[3]),
        "exact_match_with_v0": int(counts[4]),
        "out_of_range": int(counts[5]),
    }
