"""CPU tests (-m "not gpu"): the C-ABI library loads, exports every symbol include/nns_b200.h
declares, validates arguments, plans launches sanely, and FAILS LOUDLY (no CPU fallback) when no
GPU is usable."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "nns_b200.h")


def declared_symbols():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(nns_b200_[a-z_A-Z0-9]+)\s*\(", text)))


def test_header_declares_the_drop_in_symbol_with_reference_signature():
    text = open(HEADER).read()
    assert re.search(r"void\s+nns_b200_cudaCall\(int k, int m, int n, float \*s_points, float \*r_points,\s*int \*\*results\);", text)
    assert "core.cu:23-29" in text and "main.cu:74" in text and "utils.h:16-26" in text


def test_library_exports_every_declared_symbol(nns):
    syms = declared_symbols()
    assert len(syms) >= 15
    for s in syms:
        assert hasattr(nns.lib, s), f"{s} declared in include/nns_b200.h but not exported"
    assert nns.lib.nns_b200_version() == 200


def test_header_constants_match_binding(nns):
    text = open(HEADER).read()
    assert int(re.search(r"#define NNS_B200_REF_BLOCK (\d+)", text).group(1)) == nns.REF_BLOCK
    assert int(re.search(r"#define NNS_B200_KEY_INIT (0x[0-9A-Fa-f]+)", text).group(1), 16) == nns.KEY_INIT
    # (+INF, 0): the float bits of +INF in the high word
    assert nns.KEY_INIT >> 32 == np.array([np.inf], np.float32).view(np.uint32)[0]


def tensor_geometry(k):
    """Mirror of tensor_geom() in csrc/tensor_search.cu: (64-column swizzled blocks, 16-column steps)."""
    ndata = 3 * k if k <= 42 else k  # split-precision BF16 columns up to k = 42
    if ndata + 3 <= 16:
        return 0, 1
    if ndata + 3 <= 32:
        return 0, 2
    if ndata + 3 <= 64:
        return 1, 0
    if ndata <= 64:
        return 1, 1
    if ndata + 3 <= 128:
        return 2, 0
    if ndata <= 128:
        return 2, 1
    return (ndata + 3 + 63) // 64, 0  # 128 < k <= 509: whole 64-column blocks, K-loop kernel (tensor_longk.cu)


def test_index_geometry(nns):
    # 32-float header + (k coordinate rows + 1 norm row) x 128 lanes per block of 128 points, then (k <= 509)
    # the tensor section: 1024-float header + per block a BF16 operand image of 128 rows x (KB * 128 + KS * 32) B
    def floats(k, n):
        if n <= 0:
            return 0
        nb = (n + 127) // 128
        total = 32 + nb * (k + 1) * 128
        if k <= 509:
            kb, ks = tensor_geometry(k)
            total += 1024 + nb * 128 * (kb * 128 + ks * 32) // 4
        return total

    assert nns.index_floats(3, 0) == 0
    assert [tensor_geometry(k) for k in (1, 3, 4, 5, 9, 10, 16, 20, 21, 41, 42, 43, 61, 62, 64, 65, 125, 126, 128, 129, 189, 190, 317, 318, 509)] == [
        (0, 1), (0, 1), (0, 1), (0, 2), (0, 2), (1, 0), (1, 0), (1, 0), (1, 1), (2, 0), (2, 1), (1, 0), (1, 0), (1, 1),
        (1, 1), (2, 0), (2, 0), (2, 1), (2, 1), (3, 0), (3, 0), (4, 0), (5, 0), (6, 0), (8, 0)]
    for k, n in [(3, 1), (3, 129), (16, 16777216), (128, 128), (40, 129), (50, 129), (129, 128), (3, 128), (22, 128),
                 (4, 1000), (5, 1000), (10, 77), (21, 4096), (64, 300), (200, 5), (509, 300), (510, 300), (1000, 5)]:
        assert nns.index_floats(k, n) == floats(k, n), (k, n)
    assert nns.index_floats(3, 1) == 32 + 4 * 128 + 1024 + 128 * 32 // 4  # k = 3: one K = 16 step per reference
    assert nns.index_floats(510, 128) == 32 + 511 * 128                   # k > 509: no tensor section
    assert nns.lib.nns_b200_workspace_bytes(3, 10, 129) >= (32 + 2 * 4 * 128) * 4 + 80


def test_argument_validation_needs_no_gpu(nns):
    out = np.zeros(4, np.int32)
    s = np.zeros((4, 3), np.float32)
    assert nns.lib.nns_b200_search_host(0, 4, 4, s.ctypes.data, s.ctypes.data, out.ctypes.data) == nns.ERR_INVALID
    assert nns.lib.nns_b200_search_host(3, -1, 4, s.ctypes.data, s.ctypes.data, out.ctypes.data) == nns.ERR_INVALID
    assert nns.lib.nns_b200_search_host(3, 4, 4, None, s.ctypes.data, out.ctypes.data) == nns.ERR_INVALID
    assert b"NULL" in nns.lib.nns_b200_last_error()
    assert nns.lib.nns_b200_search_host(3, 0, 4, s.ctypes.data, s.ctypes.data, out.ctypes.data) == nns.OK  # nothing to do
    assert nns.lib.nns_b200_search_multi(3, 4, 4, s.ctypes.data, s.ctypes.data, out.ctypes.data, 1, 7) == nns.ERR_INVALID
    assert nns.lib.nns_b200_search_host_dist(3, 4, 4, s.ctypes.data, s.ctypes.data, out.ctypes.data, None) == nns.ERR_INVALID
    with pytest.raises(nns.NnsError):
        nns.plan(0, 1, 1)


def test_extension_entry_points_validate_before_touching_the_gpu(nns):
    import ctypes

    s = np.zeros((4, 3), np.float32)
    out = np.zeros((4, 4), np.int32)
    # K nearest neighbours: K must be 1..32, k <= 1024
    assert nns.lib.nns_b200_search_topk_host(3, 4, 4, 0, s.ctypes.data, s.ctypes.data, out.ctypes.data, None) == nns.ERR_UNSUPPORTED
    assert nns.lib.nns_b200_search_topk_host(3, 4, 4, 33, s.ctypes.data, s.ctypes.data, out.ctypes.data, None) == nns.ERR_UNSUPPORTED
    assert nns.lib.nns_b200_search_topk_host(0, 4, 4, 4, s.ctypes.data, s.ctypes.data, out.ctypes.data, None) == nns.ERR_INVALID
    assert nns.lib.nns_b200_search_topk_host(3, 0, 4, 4, s.ctypes.data, s.ctypes.data, out.ctypes.data, None) == nns.OK
    # tree: k <= 32
    h = ctypes.c_void_p()
    assert nns.lib.nns_b200_tree_create(0, 4, s.ctypes.data, -1, ctypes.byref(h)) == nns.ERR_INVALID
    assert nns.lib.nns_b200_tree_create(33, 4, s.ctypes.data, -1, ctypes.byref(h)) == nns.ERR_UNSUPPORTED
    assert nns.lib.nns_b200_tree_create(3, 4, None, -1, ctypes.byref(h)) == nns.ERR_INVALID
    assert nns.lib.nns_b200_tree_search(None, 4, s.ctypes.data, out.ctypes.data, None) == nns.ERR_INVALID
    assert nns.lib.nns_b200_tree_destroy(None) == nns.OK
    # handles and parts
    assert nns.lib.nns_b200_index_create(0, 4, s.ctypes.data, -1, ctypes.byref(h)) == nns.ERR_INVALID
    assert nns.lib.nns_b200_index_search(None, 4, s.ctypes.data, out.ctypes.data, None) == nns.ERR_INVALID
    assert nns.lib.nns_b200_index_destroy(None) == nns.OK
    c = np.zeros(3, np.float32)
    assert nns.lib.nns_b200_sample_centre(3, 4, s.ctypes.data, c.ctypes.data) == nns.OK
    assert nns.lib.nns_b200_sample_centre(600, 4, s.ctypes.data, c.ctypes.data) == nns.ERR_INVALID
    rg = (ctypes.c_size_t * 6)()
    assert nns.lib.nns_b200_index_part_ranges(3, 1024, 128, 4, 1, rg) == nns.OK
    assert rg[0] == (32 + 1 * 4 * 128) * 4 and rg[1] == 4 * 4 * 128 * 4  # blocks of part 1 start after block 0; 4 blocks
    assert rg[4] == (8 + 1) * 4                                           # its slot in the index header
    assert nns.lib.nns_b200_index_part_ranges(3, 1024, 100, 4, 1, rg) == nns.ERR_INVALID  # slices are whole blocks


def test_sample_centre_is_the_mean_of_strided_rows(nns):
    r = np.arange(10000 * 3, dtype=np.float32).reshape(10000, 3)
    c = nns.sample_centre(3, 10000, r)
    rows = r[(np.arange(4096, dtype=np.int64) * 10000) // 4096]
    np.testing.assert_allclose(c, rows.astype(np.float64).mean(0), rtol=1e-6)
    r[5, 1] = np.nan  # non-finite coordinates do not steer the centre
    assert np.isfinite(nns.sample_centre(3, 10000, r)).all()


def test_no_cpu_fallback_fails_loudly_without_gpu(nns):
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present; the failure path is exercised on the CPU container")
    s = np.zeros((4, 3), np.float32)
    with pytest.raises(nns.NnsError) as e:
        nns.search_host(3, 4, 4, s, s)
    assert e.value.status == nns.ERR_CUDA
    assert nns.launch_count() == 0


def test_plan_paths(nns):
    # BASELINE config C2: large enough for the split-precision tcgen05 screen (strips of 256 queries) ...
    p = nns.plan(3, 65536, 4194304)
    assert p["path"] == 2 and p["query_blocks"] == 256
    assert nns.plan(16, 262144, 16777216)["path"] == 2  # ... and so is C3
    assert nns.plan(3, 1024, 65536)["path"] == 0        # C1 is not: register-blocked FP32 kernel
    assert nns.plan(3, 65536, 4194304, nns.FLAG_EXACT_FORM)["path"] == 0  # V0's formulation is FP32-only
    p = nns.plan(3, 65536, 4194304, nns.FLAG_FORCE_LOWK)  # the FP32 kernel for the same shape
    assert p["path"] == 0 and p["q"] in (4, 8) and p["warps"] == 8
    assert p["splits"] * p["blocks_per_split"] >= 4194304 // 128
    assert (p["splits"] - 1) * p["blocks_per_split"] < 4194304 // 128  # no empty split
    assert p["blocks_per_split"] % 8 == 0  # whole tiles (k=3: 8 blocks per tile)
    assert p["smem"] <= 227 * 1024
    assert nns.plan(128, 1024, 65536)["path"] == 2  # 32 < k <= 128, m >= 256 -> tcgen05 path
    assert nns.plan(128, 100, 65536)["path"] == 1  # few queries -> reference-parallel FP32 kernel
    assert nns.plan(200, 1024, 65536)["path"] == 2  # 128 < k <= 509 -> tcgen05 K-loop kernel
    assert nns.plan(600, 1024, 65536)["path"] == 1  # k > 509 -> reference-parallel FP32 kernel
    assert nns.plan(128, 1024, 65536, nns.FLAG_EXACT_FORM)["path"] == 1  # V0's formulation on every pair: never the screen
    assert nns.plan(3, 1, 65536)["path"] == 1  # the reference's m = 1 shapes are reference-parallel
    assert nns.plan(3, 1, 65536, nns.FLAG_FORCE_LOWK)["path"] == 0
    assert nns.plan(3, 4096, 65536, nns.FLAG_FORCE_WIDE)["path"] == 1
    with pytest.raises(nns.NnsError):
        nns.plan(33, 1024, 1024, nns.FLAG_FORCE_LOWK)


@pytest.mark.parametrize("k", [1, 2, 3, 4, 7, 8, 16, 17, 32])
@pytest.mark.parametrize("m,n", [(16, 1), (1000, 127), (1024, 65536), (65536, 4194304), (300000, 1000), (16777216, 16777216)])
def test_plan_covers_every_reference_block_and_query(nns, k, m, n):
    for q_over in (0,):
        p = nns.plan(k, m, n, nns.flag_overrides(q=q_over) | nns.FLAG_FORCE_LOWK)
        nblocks = (n + 127) // 128
        assert p["path"] == 0
        assert p["query_blocks"] * 32 * p["warps"] * p["q"] >= m
        assert (p["query_blocks"] - 1) * 32 * p["warps"] * p["q"] < m
        assert p["splits"] * p["blocks_per_split"] >= nblocks
        assert (p["splits"] - 1) * p["blocks_per_split"] < max(nblocks, 1)
        assert 1 <= p["splits"] <= 65535 and p["query_blocks"] >= 1


def test_plan_overrides(nns):
    p = nns.plan(3, 65536, 4194304, nns.flag_overrides(q=8, warps=4, stages=3) | nns.FLAG_FORCE_LOWK)
    assert (p["q"], p["warps"], p["stages"]) == (8, 4, 3)
    with pytest.raises(nns.NnsError):
        nns.plan(3, 65536, 4194304, nns.flag_overrides(q=5) | nns.FLAG_FORCE_LOWK)


def test_tensor_bound_is_a_pure_host_function(nns):
    """nns_b200_tensor_bound needs no GPU: geometry, scales and E(q) of the tcgen05 screen for given norms"""
    b = nns.tensor_bound(16, 0, 1.0, 1.2, 1.2)
    assert b["kp"] == 64 and b["s"] == 1.0 and b["t"] == 1.0 and 0 < b["E"] < 1e-3      # split-precision BF16
    f = nns.tensor_bound(16, 2, 1.0, 1.2, 1.2)
    assert f["kp"] == 32 and f["s"] == 4.0 and f["t"] > 0 and b["E"] < f["E"] < 2e-2      # plain F16: a wider band, half the columns
    assert nns.tensor_bound(128, 0, 3.0, 3.3, 3.3)["kp"] == 144
    assert nns.tensor_bound(64, 2, 1e9, 1.0, 1.0)["t"] == 0.0                              # a query 1e9 radii out cannot be screened
    assert nns.tensor_bound(64, 2, 1.0, 500.0, 1.0)["t"] == 0.0                            # nor can any, if the sample missed the radius 500x
    for bad in ((0, 0), (510, 0), (9, 2), (129, 2), (16, 1)):
        rc = nns.lib.nns_b200_tensor_bound(bad[0], bad[1], 1.0, 1.0, 1.0, (__import__("ctypes").c_float * 4)())
        assert rc == nns.ERR_INVALID if hasattr(nns, "ERR_INVALID") else rc == 1


def test_screen_geometry_of_every_k_matches_the_python_mirror(nns):
    """contraction columns of the operand images for every k, both modes: the library (nns_b200_tensor_bound) against the
    mirrors the CPU emulation and bench.py use (tests/test_tensor_bound.geometry, nns_b200.tensor_kp)"""
    from test_tensor_bound import geometry

    for k in range(1, 510):
        assert nns.tensor_bound(k, 0, 1.0, 1.0, 1.0)["kp"] == geometry(k)[0] == nns.tensor_kp(k), k
        if 10 <= k <= 128:
            assert nns.tensor_bound(k, 2, 1.0, 1.0, 1.0)["kp"] == geometry(k, True)[0], k
