"""GPU parity tests (-m gpu): the CUDA path, called through the C ABI (ctypes -> libnns_b200.so),
against the V0 oracle on the same seeded inputs, against the committed golden vectors produced by
the reference's own V0, and -- at BASELINE.json's full sizes -- through size-independent
properties.  Bars: with NNS_B200_FLAG_V0_ROUNDING the indices are IDENTICAL to V0 (same FP32
rounding, lowest index on ties); in the default FMA mode they satisfy the north-star rule
(FP64 distance within 1e-5 relative of the minimum, exact ties -> lowest index), which on
grid-snapped data again means identical."""
import os

import numpy as np
import pytest

from conftest import make_case
from test_oracle import golden_cases

pytestmark = pytest.mark.gpu
REL_TOL = 1e-5  # north_star: indices identical wherever the best two distances differ by > 1e-5 relative


@pytest.fixture(scope="module")
def torch_mod():
    import torch

    assert torch.cuda.is_available()
    return torch


def dev(torch, a):
    return torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32)).cuda()


def gpu_search(nns, torch, s, r, flags=0):
    idx = nns.DeviceIndex(dev(torch, r)).search(dev(torch, s), flags)
    torch.cuda.synchronize()
    return idx.cpu().numpy()


def assert_rule(oracle, k, m, n, s, r, g, v, exact_expected):
    if exact_expected:
        assert np.array_equal(g, v), f"{int((g != v).sum())} of {m} indices differ from V0"
    else:
        rep = oracle.check_tie_rule(k, m, n, s, r, g, v, REL_TOL)
        assert rep["violations"] == 0 and rep["oracle_anomalies"] == 0, rep
        assert rep["exact_match_with_v0"] >= m - max(2, m // 500), rep  # SURVEY hard part 3: flips are ~1e-5 rare


def test_known_answers_through_the_drop_in_symbol(nns):
    r = np.array([[5, 5, 5], [1, 1, 1], [1, 1, 1], [1, 1, 1], [9, 9, 9]], np.float32)
    q = np.array([[1, 1, 1], [0, 0, 0], [np.nan, 0, 0], [np.inf, 0, 0], [9, 9, 8]], np.float32)
    assert nns.cudaCall(3, 5, 5, q, r).tolist() == [1, 1, 0, 0, 4]
    assert nns.cudaCall(3, 5, 0, q, r[:0]).tolist() == [0, 0, 0, 0, 0]
    r2 = np.array([[np.nan, 0, 0], [2, 2, 2], [0.5, 0.5, 0.5]], np.float32)
    assert nns.cudaCall(3, 2, 3, q[:2], r2).tolist() == [2, 2]
    assert nns.cudaCall(3, 0, 5, q[:0], r).shape == (0,)
    assert nns.launch_count() > 0


@pytest.mark.parametrize("path", ["auto", "lowk", "wide"])
@pytest.mark.parametrize("rounding", ["v0", "fma", "filter"])
def test_golden_vectors(nns, oracle, torch_mod, path, rounding):
    flags = {"auto": 0, "lowk": nns.FLAG_FORCE_LOWK, "wide": nns.FLAG_FORCE_WIDE}[path]
    if rounding == "v0":
        flags |= nns.FLAG_V0_ROUNDING
    elif rounding == "fma":
        flags |= nns.FLAG_EXACT_FORM
    ran = 0
    for kind, k, m, n, seed, idx, _ in golden_cases():
        if path == "lowk" and k > 32:
            continue
        if path == "wide" and m * n > 3e7:
            continue
        s, r = make_case(kind, k, m, n, seed)
        g = gpu_search(nns, torch_mod, s, r, flags)
        assert_rule(oracle, k, m, n, s, r, g, idx, rounding == "v0" or kind in ("grid", "clustered"))
        ran += 1
    assert ran >= 15


def test_config_c1_through_host_abi_is_identical_to_v0(nns, oracle):
    # BASELINE config C1: k=3, m=1024, n=65536 -- the correctness gate, via the drop-in symbol
    k, m, n = 3, 1024, 65536
    s, r = make_case("uniform", k, m, n, 1000)
    v = oracle.v0(k, m, n, s, r)
    g = nns.cudaCall(k, m, n, s, r)
    rep = oracle.check_tie_rule(k, m, n, s, r, g, v, REL_TOL)
    assert rep["violations"] == 0, rep
    assert rep["exact_match_with_v0"] == m, rep  # 100% agreement on C1
    # and on the data the reference's own generator makes for its first (3, 1, 1024) sample
    from nns_b200 import datagen

    s2, r2 = datagen.reference_rand_sample(3, 1, 1024, 1000)
    assert np.array_equal(nns.cudaCall(3, 1, 1024, s2, r2), oracle.v0(3, 1, 1024, s2, r2))


SWEEP = [(k, m, n) for k in (1, 2, 3, 4, 5, 6, 8, 12, 16, 17, 24, 31, 32) for (m, n) in ((16, 1), (100, 127), (257, 1000), (1500, 5000))]


@pytest.mark.parametrize("k,m,n", SWEEP)
def test_lowk_shape_sweep_identical_with_v0_rounding(nns, oracle, torch_mod, k, m, n):
    s, r = make_case("uniform", k, m, n, 100 + k)
    v = oracle.v0(k, m, n, s, r)
    g = gpu_search(nns, torch_mod, s, r, nns.FLAG_FORCE_LOWK | nns.FLAG_V0_ROUNDING)
    assert np.array_equal(g, v)
    from nns_b200 import nns_plan_q

    ref = None
    for q in nns_plan_q(k):  # both register blockings, FMA modes: north-star rule
        for form in (nns.FLAG_EXACT_FORM, 0):  # every pair in V0 form / norm-expansion screen + exact survivors
            g = gpu_search(nns, torch_mod, s, r, nns.FLAG_FORCE_LOWK | form | nns.flag_overrides(q=q))
            assert_rule(oracle, k, m, n, s, r, g, v, False)
            ref = g if ref is None else ref
            assert np.array_equal(g, ref)  # the screen never changes the answer


@pytest.mark.parametrize("k", [3, 16])
@pytest.mark.parametrize("warps", [1, 2, 4, 8])
def test_lowk_every_cta_geometry(nns, oracle, torch_mod, k, warps):
    m, n = 3000, 20000
    s, r = make_case("grid", k, m, n, 7)
    v = oracle.v0(k, m, n, s, r)
    from nns_b200 import nns_plan_q

    for q in nns_plan_q(k):
        for stages in (2, 4):
            for form in (0, nns.FLAG_EXACT_FORM):
                g = gpu_search(nns, torch_mod, s, r, nns.FLAG_FORCE_LOWK | form | nns.flag_overrides(q=q, warps=warps, stages=stages))
                assert np.array_equal(g, v), (k, warps, q, stages, form)  # grid data: exact in FMA mode too


@pytest.mark.parametrize("k,m,n", [(33, 20, 900), (64, 37, 3000), (128, 64, 4096), (200, 5, 1000), (3, 1, 70000), (16, 3, 33000), (3, 7, 129)])
def test_wide_path(nns, oracle, torch_mod, k, m, n):
    s, r = make_case("uniform", k, m, n, 55)
    v = oracle.v0(k, m, n, s, r)
    assert np.array_equal(gpu_search(nns, torch_mod, s, r, nns.FLAG_FORCE_WIDE | nns.FLAG_V0_ROUNDING), v)
    assert_rule(oracle, k, m, n, s, r, gpu_search(nns, torch_mod, s, r, nns.FLAG_FORCE_WIDE), v, False)


def test_special_values(nns, oracle, torch_mod):
    k, m, n = 3, 64, 1000
    s, r = make_case("uniform", k, m, n, 9)
    s, r = s.copy(), r.copy()
    s[1] = np.nan
    s[2, 1] = np.inf
    s[3] = 1e30  # distances overflow to +INF in FP32 -> never '<' INF -> index 0
    r[0] = np.nan
    r[5, 2] = np.inf
    r[17] = -np.inf
    v = oracle.v0(k, m, n, s, r)
    assert v[1] == 0 and v[2] == 0 and v[3] == 0
    for flags in (nns.FLAG_FORCE_LOWK, nns.FLAG_FORCE_LOWK | nns.FLAG_EXACT_FORM, nns.FLAG_FORCE_WIDE,
                  nns.FLAG_FORCE_LOWK | nns.FLAG_V0_ROUNDING):
        g = gpu_search(nns, torch_mod, s, r, flags)
        assert np.array_equal(g, v), flags


@pytest.mark.parametrize("k", [1, 3, 16, 32])
@pytest.mark.parametrize("case", ["offset1000", "offset1e6", "tiny", "huge1e15", "huge1e25", "denormal", "mixed_scale", "nan_refs"])
def test_filter_adversarial_magnitudes(nns, oracle, torch_mod, k, case):
    """The norm-expansion screen must never change the answer: data built to stress its error bound
    (large common offsets => catastrophic cancellation in |r|^2 - 2q.r, magnitudes that overflow
    FP32 norms, denormals, mixed scales, NaN reference points).  The screened kernel must return
    exactly what the exact-form kernel returns, which must satisfy the rule against V0."""
    m, n = 600, 20000
    s, r = make_case("uniform", k, m, n, 31)
    s, r = s.astype(np.float64), r.astype(np.float64)
    if case == "offset1000":
        s, r = s + 1000.0, r + 1000.0
    elif case == "offset1e6":
        s, r = s * 4 + 1e6, r * 4 + 1e6
    elif case == "tiny":
        s, r = s * 1e-18, r * 1e-18
    elif case == "huge1e15":
        s, r = s * 1e15, r * 1e15
    elif case == "huge1e25":  # |r|^2 overflows FP32: the screen must disable itself
        s, r = s * 1e25, r * 1e25
    elif case == "denormal":
        s, r = s * 1e-42, r * 1e-42
    elif case == "mixed_scale":
        r[::7] *= 1e6
        s[::5] *= 1e-4
    s, r = s.astype(np.float32), r.astype(np.float32)
    if case == "nan_refs":
        r[::11] = np.nan
        r[5, min(1, k - 1)] = np.inf
    v = oracle.v0(k, m, n, s, r)
    ge = gpu_search(nns, torch_mod, s, r, nns.FLAG_FORCE_LOWK | nns.FLAG_EXACT_FORM)
    gf = gpu_search(nns, torch_mod, s, r, nns.FLAG_FORCE_LOWK)
    assert np.array_equal(gf, ge), int((gf != ge).sum())
    gv = gpu_search(nns, torch_mod, s, r, nns.FLAG_FORCE_LOWK | nns.FLAG_V0_ROUNDING)
    assert np.array_equal(gv, v)
    # where V0's own FP32 distances overflow or lose bits to underflow (huge1e25, denormal, k = 1 at 1e-18
    # scale) the FP64 rule does not describe V0 any more; the V0 equality above covers those cases
    with np.errstate(all="ignore"):
        d_v0 = ((s - r[v]) ** 2).sum(axis=1, dtype=np.float32)
    if case != "huge1e25" and not (d_v0 < 1e-30).any():
        rep = oracle.check_tie_rule(k, m, n, s, r, gf, v, REL_TOL)
        assert rep["outside_band"] == 0 and rep["out_of_range"] == 0, rep


def test_duplicates_resolve_to_lowest_index_everywhere(nns, oracle, torch_mod):
    # clustered + grid-snapped + duplicated points (BASELINE config C5's construction, reduced)
    k, m, n = 3, 8192, 200000
    s, r = make_case("clustered", k, m, n, 1000)
    v, _ = oracle.v0_omp(k, m, n, s, r)
    for flags in (0, nns.flag_overrides(q=8), nns.FLAG_EXACT_FORM, nns.FLAG_FORCE_WIDE):
        if flags == nns.FLAG_FORCE_WIDE:
            g = gpu_search(nns, torch_mod, s[:512], r, flags)
            assert np.array_equal(g, v[:512])
        else:
            g = gpu_search(nns, torch_mod, s, r, flags)
            assert np.array_equal(g, v), int((g != v).sum())


def test_reference_shard_count_invariance(nns, oracle, torch_mod):
    # emulates the reference-sharded multi-GPU path on one GPU: every shard accumulates into the
    # same packed keys with its own index base; result must not depend on the shard count
    torch = torch_mod
    k, m, n = 16, 2048, 50000
    s, r = make_case("grid", k, m, n, 3)
    v, _ = oracle.v0_omp(k, m, n, s, r)
    dq = dev(torch, s)
    results = []
    for G in (1, 2, 3, 4, 8):
        blocks = (n + 127) // 128
        per = ((blocks + G - 1) // G) * 128
        keys = None
        for g in reversed(range(G)):  # any order
            r0 = g * per
            if r0 >= n:
                continue
            idxobj = nns.DeviceIndex(dev(torch, r[r0:r0 + per]), index_base=r0)
            keys = idxobj.new_keys(m) if keys is None else keys
            idxobj.search_keys(dq, keys)
        out = nns.unpack_keys(keys, m).cpu().numpy()
        results.append(out)
        assert np.array_equal(out, v), G
    assert all(np.array_equal(results[0], x) for x in results)


def test_host_ingest_chunking_and_search_multi(nns, oracle):
    # n large enough for >= 2 ingest chunks (32 MiB of AoS each) through the host-pointer ABI
    k, m, n = 3, 200, 3_000_000
    s, r = make_case("uniform", k, m, n, 77)
    v, _ = oracle.v0_omp(k, m, n, s, r)
    g = nns.search_host(k, m, n, s, r)
    rep = oracle.check_tie_rule(k, m, n, s, r, g, v, REL_TOL)
    assert rep["violations"] == 0 and rep["exact_match_with_v0"] >= m - 1, rep
    for mode in (0, 1):
        g2 = nns.search_multi(k, m, n, s, r, num_gpus=1, shard_mode=mode)
        assert np.array_equal(g2, g)


@pytest.mark.parametrize("k,m,n", [(3, 700, 30000), (16, 300, 9000), (128, 512, 8192), (200, 9, 3000), (3, 4, 0)])
def test_search_host_dist_returns_the_fp32_distance_of_the_reported_neighbour(nns, oracle, k, m, n):
    """The distances-out extension (the 1-NN part): indices as nns_b200_search_host,
    and the FP32 distance of that neighbour accumulated like the engine does (ascending dimensions,
    FMA) -- equal to an FP64 evaluation to FP32 round-off; +INF when there is no reference."""
    s, r = make_case("uniform", k, m, max(n, 1), 5)
    r = r[:n]
    idx, dist = nns.search_host_dist(k, m, n, s, r)
    assert np.array_equal(idx, nns.search_host(k, m, n, s, r))
    if n == 0:
        assert np.all(idx == 0) and np.all(np.isinf(dist))
        return
    d64 = ((s.astype(np.float64) - r[idx].astype(np.float64)) ** 2).sum(axis=1)
    assert np.allclose(dist, d64, rtol=(k + 4) * 2.0 ** -23, atol=0.0)
    v = oracle.v0(k, m, n, s, r)
    rep = oracle.check_tie_rule(k, m, n, s, r, idx, v, REL_TOL)
    assert rep["violations"] == 0, rep


def test_search_device_one_shot(nns, oracle, torch_mod):
    k, m, n = 5, 700, 9000
    s, r = make_case("uniform", k, m, n, 12)
    v = oracle.v0(k, m, n, s, r)
    g = nns.search_device(dev(torch_mod, s), dev(torch_mod, r), nns.FLAG_V0_ROUNDING).cpu().numpy()
    assert np.array_equal(g, v)


def test_full_size_c2_properties(nns, oracle, torch_mod):
    # BASELINE config C2 at full size (k=3, m=65,536, n=4,194,304): the oracle cannot finish all of
    # it, so: (a) a seeded sample of queries is checked against V0 over the FULL reference set,
    # (b) two different register blockings agree bit-for-bit on keys (checksum of all keys),
    # (c) idempotence, (d) reversing the reference order maps idx -> n-1-idx (no exact ties in
    # uniform data, so the argmin is order-independent).
    torch = torch_mod
    k, m, n = 3, 65536, 4194304
    s, r = make_case("uniform", k, m, n, 1000)
    dq, dr = dev(torch, s), dev(torch, r)
    index = nns.DeviceIndex(dr)
    keys8 = index.search_keys(dq, index.new_keys(m), nns.FLAG_FORCE_LOWK | nns.flag_overrides(q=8))
    keys4 = index.search_keys(dq, index.new_keys(m), nns.FLAG_FORCE_LOWK)
    keys_exact = index.search_keys(dq, index.new_keys(m), nns.FLAG_EXACT_FORM)
    assert nns.plan(k, m, n)["path"] == 2  # the planner sends this shape to the split-precision tcgen05 screen
    keys_t = index.search_keys(dq, index.new_keys(m))
    torch.cuda.synchronize()
    st = nns.tensor_stats()
    assert st["overflow"] == 0 and st["candidates"] < 64 * m, st
    assert torch.equal(keys8, keys4)
    assert torch.equal(keys8, keys_exact)  # screened and exact-form kernels: bit-identical (dist, idx) keys
    assert torch.equal(keys8, keys_t)      # ... and so is the tensor-core screen + exact FP32 re-score
    assert int(keys8.sum().item()) == int(keys4.sum().item())
    again = index.search_keys(dq, keys8.clone())  # idempotent: min with itself
    assert torch.equal(again, keys8)
    g = nns.unpack_keys(keys8, m).cpu().numpy()
    sample = np.random.default_rng(1000).permutation(m)[:256]
    v, _ = oracle.v0_omp(k, 256, n, s[sample], r)
    rep = oracle.check_tie_rule(k, 256, n, s[sample], r, g[sample], v, REL_TOL)
    assert rep["violations"] == 0 and rep["exact_match_with_v0"] >= 255, rep
    grev = nns.DeviceIndex(torch.flip(dr, dims=[0]).contiguous()).search(dq).cpu().numpy()
    assert np.array_equal(n - 1 - grev, g)


@pytest.mark.parametrize("name,k,m,n,nsample", [("C3", 16, 262144, 16777216, 48), ("C4", 128, 1048576, 1048576, 96)])
def test_full_size_c3_c4_properties(nns, oracle, torch_mod, name, k, m, n, nsample):
    """BASELINE configs C3 (k=16, m=262,144, n=16,777,216) and C4 (k=128, m=n=1,048,576) at full size
    on the tcgen05 path the planner picks for them: (a) a seeded sample of queries against V0 over the
    FULL reference set, (b) idempotence of the packed-key minimum, (c) searching the two halves of
    the reference set separately (index_base) and min-merging the keys equals the one-shot search
    bit for bit, (d) C3: the FP32 screened kernel returns the same (dist, idx) keys, (e) the screen
    stayed selective (no overflow, a few dozen candidates per query)."""
    torch = torch_mod
    s, r = make_case("uniform", k, m, n, 1000)
    dq, dr = dev(torch, s), dev(torch, r)
    index = nns.DeviceIndex(dr)
    assert nns.plan(k, m, n)["path"] == 2
    keys = index.search_keys(dq, index.new_keys(m))
    torch.cuda.synchronize()
    st = nns.tensor_stats()
    assert st["overflow"] == 0 and st["candidates"] < 64 * m, st
    again = index.search_keys(dq, keys.clone())
    assert torch.equal(again, keys)
    half = (n // 2 // 128) * 128
    lo, hi = nns.DeviceIndex(dr[:half].contiguous()), nns.DeviceIndex(dr[half:].contiguous(), index_base=half)
    merged = hi.search_keys(dq, lo.search_keys(dq, index.new_keys(m)))
    assert torch.equal(merged, keys)
    if k <= 32:
        assert torch.equal(index.search_keys(dq, index.new_keys(m), nns.FLAG_FORCE_LOWK), keys)
    g = nns.unpack_keys(keys, m).cpu().numpy()
    sample = np.random.default_rng(1000).permutation(m)[:nsample]
    v, _ = oracle.v0_omp(k, nsample, n, s[sample], r)
    rep = oracle.check_tie_rule(k, nsample, n, s[sample], r, g[sample], v, REL_TOL)
    assert rep["violations"] == 0 and rep["exact_match_with_v0"] >= nsample - 1, rep


def test_search_multi_on_every_visible_gpu(nns, oracle, torch_mod):
    """nns_b200_search_multi (one process, host thread per GPU): both shardings return V0's answer
    for every GPU count; needs >= 2 visible GPUs to be more than a smoke test."""
    ngpu = torch_mod.cuda.device_count()
    k, m, n = 3, 5000, 300000
    s, r = make_case("clustered", k, m, n, 1000)  # duplicates: ties must resolve to the lowest index
    v, _ = oracle.v0_omp(k, m, n, s, r)
    for g in sorted({1, min(2, ngpu), ngpu}):
        for mode in (0, 1):
            out = nns.search_multi(k, m, n, s, r, num_gpus=g, shard_mode=mode)
            assert np.array_equal(out, v), (g, mode, int((out != v).sum()))


# ---- tcgen05 path (32 < k <= 128) -------------------------------------------------------------
@pytest.mark.parametrize("k,m,n", [(128, 256, 4096), (128, 1000, 20000), (64, 700, 9000), (33, 300, 5000), (100, 513, 12345),
                                    (128, 2048, 65536), (48, 1, 1000), (128, 5, 129),
                                    # split-precision BF16 (k <= 42): the same tensor screen for low k
                                    (3, 5000, 300000), (1, 700, 20000), (2, 300, 999), (16, 2000, 50000), (21, 512, 30000),
                                    (22, 512, 30000), (32, 1024, 40000), (42, 300, 8000), (43, 300, 8000), (3, 65536, 1048576),
                                    (4, 999, 77777), (5, 600, 50000), (9, 700, 60000), (10, 700, 60000)])
def test_tensor_path_matches_v0(nns, oracle, torch_mod, k, m, n):
    s, r = make_case("uniform", k, m, n, 77)
    if m * n <= 2e9:
        v, _ = oracle.v0_omp(k, m, n, s, r)
    else:  # too big for the CPU oracle in a test: the exact-form FP32 kernel in V0 rounding stands in
        v = gpu_search(nns, torch_mod, s, r, nns.FLAG_FORCE_LOWK | nns.FLAG_V0_ROUNDING)
        sample = np.arange(0, m, m // 256)
        assert np.array_equal(v[sample], oracle.v0_omp(k, len(sample), n, s[sample], r)[0])
    g = gpu_search(nns, torch_mod, s, r, nns.FLAG_FORCE_TENSOR | nns.FLAG_V0_ROUNDING)
    assert np.array_equal(g, v), int((g != v).sum())  # exact re-score in V0 rounding: identical
    g = gpu_search(nns, torch_mod, s, r, nns.FLAG_FORCE_TENSOR)
    if m * n <= 2e9:
        assert_rule(oracle, k, m, n, s, r, g, v, False)
    st = nns.tensor_stats()
    # the tcgen05 screen itself must be selective (not rescued by the overflow fallback): a handful
    # of candidate tiles per query (running-minimum records + the 2E band), never the whole grid
    ntiles = (n + 127) // 128
    # (each of the <= ~300 reference splits of a strip emits its first tile, then records + band)
    assert st["overflow"] == 0 and 0 < st["candidates"] <= m * min(4 * ntiles, 700), st
    if ntiles >= 400 and m >= 2048:
        assert st["candidates"] <= 0.15 * m * 4 * ntiles, st  # of the m x (4 units per tile) grid
    w = gpu_search(nns, torch_mod, s, r, nns.FLAG_FORCE_WIDE if m * n <= 2e9 else nns.FLAG_FORCE_LOWK | nns.FLAG_EXACT_FORM)
    assert np.array_equal(g, w)  # same FP32 arithmetic decides in both paths


@pytest.mark.parametrize("k", [128, 3, 16])
@pytest.mark.parametrize("case", ["grid", "duplicates", "offset", "nan_inf", "all_identical", "scaled"])
def test_tensor_path_adversarial(nns, oracle, torch_mod, case, k):
    m, n = 600, 30000
    s, r = make_case("uniform", k, m, n, 5)
    s, r = s.copy(), r.copy()
    if case == "grid":  # coarse grid: many exact ties across tiles -> lowest index must win
        s, r = (np.floor(s * 2) / 2).astype(np.float32), (np.floor(r * 2) / 2).astype(np.float32)
    elif case == "duplicates":
        r[1000:2000] = r[20000:21000]
        s[::2] = r[(np.arange(0, m, 2) * 37) % n]
    elif case == "offset":
        s, r = s + 100.0, r + 100.0
    elif case == "nan_inf":
        r[::13] = np.nan
        r[7, min(3, k - 1)] = np.inf
        s[3] = np.nan
        s[4, 0] = np.inf
    elif case == "all_identical":  # every tile ties with every other
        r[:] = r[0]
    elif case == "scaled":
        s, r = s * 1e8, r * 1e8
    v, _ = oracle.v0_omp(k, m, n, s, r)
    g = gpu_search(nns, torch_mod, s, r, nns.FLAG_FORCE_TENSOR | nns.FLAG_V0_ROUNDING)
    assert np.array_equal(g, v), (case, int((g != v).sum()))
    g = gpu_search(nns, torch_mod, s, r, nns.FLAG_FORCE_TENSOR)
    st = nns.tensor_stats()
    if case not in ("all_identical", "grid"):  # massive exact ties: every unit qualifies, the buffer may overflow (the fallback is exact too)
        assert st["overflow"] == 0, (case, st)
    w = gpu_search(nns, torch_mod, s, r, nns.FLAG_FORCE_WIDE)
    assert np.array_equal(g, w), (case, int((g != w).sum()))
    # candidate-buffer overflow: a device flag makes the FP32 wide kernel redo the search
    g2 = gpu_search(nns, torch_mod, s, r, nns.FLAG_FORCE_TENSOR | nns.FLAG_TEST_TINY_CANDIDATES)
    assert nns.tensor_stats()["overflow"] == 1
    assert np.array_equal(g2, w), (case, int((g2 != w).sum()))


def test_tensor_path_through_host_abi_and_shards(nns, oracle, torch_mod):
    k, m, n = 128, 1024, 40000
    s, r = make_case("uniform", k, m, n, 9)
    v, _ = oracle.v0_omp(k, m, n, s, r)
    g = nns.cudaCall(k, m, n, s, r)  # auto path: k = 128, m >= 256 -> tensor
    rep = oracle.check_tie_rule(k, m, n, s, r, g, v, REL_TOL)
    assert rep["violations"] == 0 and rep["exact_match_with_v0"] >= m - 2, rep
    assert nns.plan(k, m, n)["path"] == 2
    # reference shards accumulate into the same keys (each shard has its own centre / image)
    torch = torch_mod
    dq = dev(torch, s)
    keys = None
    for r0, r1 in ((20096, n), (0, 20096)):
        idx = nns.DeviceIndex(dev(torch, r[r0:r1]), index_base=r0)
        keys = idx.new_keys(m) if keys is None else keys
        idx.search_keys(dq, keys, nns.FLAG_V0_ROUNDING)
    assert np.array_equal(nns.unpack_keys(keys, m).cpu().numpy(), v)


def test_c5_construction_at_one_million_points(nns, oracle, torch_mod):
    """BASELINE config C5's data (clustered Gaussians snapped to a 2^-10 grid, duplicated references,
    every 2nd query an exact copy of a reference) at m = n = 2^20: a seeded sample of 1,024 queries
    (512 of them duplicated points) must equal V0 exactly over the full reference set -- on this
    grid every FP32 operation is exact, so FMA contraction cannot change a single distance."""
    torch = torch_mod
    k, m, n = 3, 1 << 20, 1 << 20
    s, r = make_case("clustered", k, m, n, 1000)
    index = nns.DeviceIndex(dev(torch, r))
    g = index.search(dev(torch, s)).cpu().numpy()
    # this shape is planned onto the tcgen05 screen; whether the screen copes with the dense clusters or
    # hands over to the FP32 kernel (candidate overflow), the answer must be the FP32 kernel's
    assert nns.plan(k, m, n)["path"] == 2
    print("C5 @ 2^20 tensor stats:", nns.tensor_stats())
    assert np.array_equal(g, index.search(dev(torch, s), nns.FLAG_FORCE_LOWK).cpu().numpy())
    sample = np.random.default_rng(5).permutation(m)[:1024]
    assert (sample % 2 == 0).sum() >= 256
    v, _ = oracle.v0_omp(k, 1024, n, s[sample], r)
    assert np.array_equal(g[sample], v), int((g[sample] != v).sum())
    # all duplicated-point queries found a zero-distance reference with the lowest index among its copies
    dup = np.arange(0, m, 2)
    assert np.array_equal(r[g[dup]], s[dup])


# ---- round 2: ingest, handles, batching, multi-GPU ---------------------------------------------
def test_pageable_staging_and_pinned_sources_agree(nns, oracle, torch_mod):
    """The host-pointer ABI takes pageable arrays (the reference passes malloc'd memory, main.cu:27-34)
    through the pinned staging ring -- > 3 slots of 8 MiB here, so slots are reused -- and pinned ones
    straight to the copy engine; both must give V0's answer."""
    torch = torch_mod
    k, m, n = 16, 700, 600_000  # 36.6 MiB of references
    s, r = make_case("uniform", k, m, n, 21)
    v, _ = oracle.v0_omp(k, m, n, s, r)
    g_pageable = nns.search_host(k, m, n, s, r)
    rep = oracle.check_tie_rule(k, m, n, s, r, g_pageable, v, REL_TOL)
    assert rep["violations"] == 0 and rep["exact_match_with_v0"] >= m - 1, rep
    sp, rp = torch.from_numpy(s).pin_memory(), torch.from_numpy(r).pin_memory()
    g_pinned = nns.search_host(k, m, n, sp.data_ptr(), rp.data_ptr())
    assert np.array_equal(g_pinned, g_pageable)


def test_chunked_ingest_on_the_tensor_path(nns, oracle):
    """A host-pointer call that is planned onto the tcgen05 screen ingests the references in chunks too
    (every chunk is an index of its own: own centre, own operand images): k = 16, 32 MiB -> 2 chunks."""
    k, m, n = 16, 4096, 500_000
    assert nns.plan(k, m, n)["path"] == 2
    s, r = make_case("uniform", k, m, n, 22)
    sample = np.random.default_rng(2).permutation(m)[:256]
    v, _ = oracle.v0_omp(k, 256, n, s[sample], r)
    g = nns.search_host(k, m, n, s, r)
    rep = oracle.check_tie_rule(k, 256, n, s[sample], r, g[sample], v, REL_TOL)
    assert rep["violations"] == 0 and rep["exact_match_with_v0"] >= 255, rep
    assert nns.tensor_stats()["overflow"] == 0
    # the same answer from the FP32 kernel (one index, no chunks)
    import torch
    idx = nns.DeviceIndex(dev(torch, r)).search(dev(torch, s), nns.FLAG_FORCE_LOWK).cpu().numpy()
    assert np.array_equal(g, idx), int((g != idx).sum())


@pytest.mark.parametrize("kind,k,m,n", [("clustered", 3, 3000, 250_000), ("uniform", 128, 600, 30_000), ("grid", 16, 2000, 70_000)])
def test_host_index_handle_build_once_query_many(nns, oracle, kind, k, m, n):
    """nns_b200_index_create / _search / _destroy: the index stays resident; repeated searches (with and
    without distances) return V0's answer."""
    s, r = make_case(kind, k, m, n, 31)
    v, _ = oracle.v0_omp(k, m, n, s, r)
    h = nns.HostIndex(k, n, r)
    for rep_i in range(3):
        if rep_i == 1:
            g, d = h.search(m, s, return_dist=True)
            e = (s.astype(np.float64) - r[g].astype(np.float64))
            np.testing.assert_allclose(d, (e * e).sum(1), rtol=1e-5, atol=1e-12)
        else:
            g = h.search(m, s)
        rep = oracle.check_tie_rule(k, m, n, s, r, g, v, REL_TOL)
        assert rep["violations"] == 0 and rep["exact_match_with_v0"] >= m - 2, rep
    # a second, smaller query batch on the same handle
    g2 = h.search(17, s[5:22])
    assert np.array_equal(g2, g[5:22])
    h.close()
    # empty index: V0 reports 0
    h0 = nns.HostIndex(k, 0, r[:0])
    assert h0.search(4, s[:4]).tolist() == [0, 0, 0, 0]
    h0.close()


def test_tensor_path_query_batches(nns, oracle, torch_mod):
    """More than 4 waves of 256-query strips are searched in batches on a bounded scratch: the keys must
    equal the FP32 kernel's for every query of every batch."""
    torch = torch_mod
    k, n = 3, 30_000
    m = 256 * (4 * nns.device_sms() + 37) + 11  # two batches, the second ragged
    s, r = make_case("clustered", k, m, n, 41)
    index = nns.DeviceIndex(dev(torch, r))
    dq = dev(torch, s)
    a = index.search(dq, nns.FLAG_FORCE_TENSOR | nns.FLAG_V0_ROUNDING).cpu().numpy()
    assert nns.tensor_stats()["overflow"] == 0
    b = index.search(dq, nns.FLAG_FORCE_LOWK | nns.FLAG_V0_ROUNDING).cpu().numpy()
    assert np.array_equal(a, b), int((a != b).sum())
    sample = np.random.default_rng(4).permutation(m)[:512]
    v, _ = oracle.v0_omp(k, 512, n, s[sample], r)
    assert np.array_equal(a[sample], v)


def test_dense_clusters_stay_on_the_tensor_screen(nns, oracle, torch_mod):
    """BASELINE config C5's construction at n = 2^22: hundreds of 32-reference units per query fall
    inside the screen's 2E band.  The candidate capacity of a batch absorbs them (no overflow, no FP32
    fallback) and the result is V0's."""
    torch = torch_mod
    k, m, n = 3, 1 << 16, 1 << 22
    s, r = make_case("clustered", k, m, n, 1000)
    index = nns.DeviceIndex(dev(torch, r))
    g = index.search(dev(torch, s)).cpu().numpy()
    st = nns.tensor_stats()
    print("C5-style @ 2^22 tensor stats:", st)
    assert nns.plan(k, m, n)["path"] == 2 and st["overflow"] == 0, st
    sample = np.random.default_rng(6).permutation(m)[:512]
    v, _ = oracle.v0_omp(k, 512, n, s[sample], r)
    assert np.array_equal(g[sample], v), int((g[sample] != v).sum())
    assert np.array_equal(g, index.search(dev(torch, s), nns.FLAG_FORCE_LOWK).cpu().numpy())


@pytest.mark.parametrize("kind,k,m,n", [("clustered", 3, 5000, 300_000), ("uniform", 16, 3000, 1_200_000), ("uniform", 128, 4096, 70_000)])
def test_search_multi_gpu_counts_and_modes(nns, oracle, torch_mod, kind, k, m, n):
    """nns_b200_search_multi on 1, 2, 4, 8 ... visible GPUs, both shardings: query-sharded uses the
    sharded ingest whose build kernels store every slice into all peers' indices (fused NVLink
    all-gather); reference-sharded folds the packed keys into GPU 0 with system-scope red.min.  Needs
    >= 2 visible GPUs to be more than a smoke test (the G = 1 legs always run)."""
    ngpu = torch_mod.cuda.device_count()
    s, r = make_case(kind, k, m, n, 1000)
    sample = np.random.default_rng(8).permutation(m)[:384]
    v, _ = oracle.v0_omp(k, 384, n, s[sample], r)
    base = None
    for g in [x for x in (1, 2, 3, 4, 8) if x <= ngpu]:
        for mode in (0, 1):
            out = nns.search_multi(k, m, n, s, r, num_gpus=g, shard_mode=mode)
            if kind == "uniform":
                rep = oracle.check_tie_rule(k, 384, n, s[sample], r, out[sample], v, REL_TOL)
                assert rep["violations"] == 0 and rep["exact_match_with_v0"] >= 383, (g, mode, rep)
            else:
                assert np.array_equal(out[sample], v), (g, mode, int((out[sample] != v).sum()))
            if base is None:
                base = out
            # identical for every GPU count and sharding on tie-free / grid-snapped data
            assert np.array_equal(out, base) or kind == "uniform" and (out != base).sum() <= 2, (g, mode, int((out != base).sum()))


def test_index_parts_and_nccl_gather_world1(nns, oracle, torch_mod):
    """nns_b200_index_build_part + exchange + nns_b200_index_finish (sharding.gather_built_index) with a
    one-rank NCCL group: the gathered index must answer like an index built in one piece.  (The world > 1
    exchange is the same code with all-gathers over NVLink; bench.py --gpus N exercises it.)"""
    torch = torch_mod
    import torch.distributed as dist

    from nns_b200 import sharding

    created = False
    if not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29533")
        dist.init_process_group("nccl", rank=0, world_size=1, device_id=torch.device("cuda", 0))
        created = True
    try:
        for (k, m, n) in [(3, 4096, 100_000), (128, 1024, 20_000)]:
            s, r = make_case("uniform", k, m, n, 51)
            v, _ = oracle.v0_omp(k, m, n, s, r)
            index = sharding.gather_built_index(k, n, r, 0, 1, torch.device("cuda", 0))
            for flags in (nns.FLAG_V0_ROUNDING, nns.FLAG_V0_ROUNDING | nns.FLAG_FORCE_TENSOR):
                g = index.search(dev(torch, s), flags).cpu().numpy()
                assert np.array_equal(g, v), (k, flags, int((g != v).sum()))
    finally:
        if created:
            dist.destroy_process_group()


# ---- K nearest neighbours (extension) ------------------------------------------
@pytest.mark.parametrize("kind,k,m,n,K", [("uniform", 3, 1024, 65536, 8), ("grid", 3, 500, 20000, 32), ("clustered", 3, 2000, 1 << 20, 16),
                                          ("uniform", 16, 300, 50000, 5), ("uniform", 128, 100, 9000, 32), ("uniform", 200, 33, 2000, 3),
                                          ("uniform", 3, 50, 20, 32), ("uniform", 3, 5, 0, 4), ("grid", 2, 1, 300000, 1)])
def test_topk_matches_the_v0_form_oracle(nns, oracle, torch_mod, kind, k, m, n, K):
    """nns_b200_topk_keys / nns_b200_search_topk_host: the K nearest references by (V0-form FP32 distance, index)
    must equal the oracle exactly with NNS_B200_FLAG_V0_ROUNDING (duplicates, grids, n < K, n = 0 included),
    and K = 1 must be V0's answer."""
    torch = torch_mod
    s, r = make_case(kind, k, m, max(n, 1), 61)
    r = r[:n]
    want_i, want_d = oracle.v0_topk(k, m, n, K, s, r)
    if n > 0:
        index = nns.DeviceIndex(dev(torch, r))
        gi, gd = index.topk(dev(torch, s), K, nns.FLAG_V0_ROUNDING)
        gi, gd = gi.cpu().numpy(), gd.cpu().numpy()
        assert np.array_equal(gi, want_i), int((gi != want_i).sum())
        assert np.array_equal(gd, want_d)
        # two reference shards accumulate into the same lists
        h = (n // 2 + 127) // 128 * 128
        if 0 < h < n:
            keys = None
            for r0, r1 in ((h, n), (0, h)):
                part = nns.DeviceIndex(dev(torch, r[r0:r1]), index_base=r0)
                keys = part.topk_keys(dev(torch, s), K, keys, nns.FLAG_V0_ROUNDING)
            assert np.array_equal((keys.cpu().numpy().astype(np.uint64) & np.uint64(0xFFFFFFFF)).astype(np.int64)[want_i >= 0], want_i[want_i >= 0])
    # host ABI (default FMA rounding): same sets up to FP32 rounding of near-ties -- on grids identical
    hi, hd = nns.search_topk_host(k, m, n, K, s, r)
    if kind != "uniform":
        assert np.array_equal(hi, want_i)
    else:
        assert (hi != want_i).mean() < 0.01
    np.testing.assert_allclose(hd[want_i >= 0], want_d[want_i >= 0], rtol=2e-6)
    assert np.all(hi[want_i < 0] == -1)
    if n > 0 and kind != "uniform":
        v, _ = oracle.v0_omp(k, m, n, s, r)
        assert np.array_equal(hi[:, 0], v)


def test_precision_mode_probe_picks_plain_or_split_per_index(nns, oracle, torch_mod):
    """For 10 <= k <= 42 the index build probes the data (2-NN distances of sample points vs the plain-F16
    error band) and builds plain F16 operand images (a third of the tensor work, 16-bit accumulators) when the
    band is selective, split-precision BF16 ones otherwise.  Uniform 16-D data -> plain (32 columns); the same points
    squeezed onto a 2-D sheet inside the 16-D cube -> split (64 columns).  Both answers must be V0's."""
    torch = torch_mod
    k, m, n = 16, 2048, 400_000
    s, r = make_case("uniform", k, m, n, 71)
    for name, want_kp, want_mode in (("uniform", 32, "f16"), ("sheet", 64, "bf16")):
        if name == "sheet":  # coordinates 2..15 are tiny multiples of the first two: intrinsic dimension 2
            rr, ss = r.copy(), s.copy()
            for t in range(2, k):
                rr[:, t] = np.float32(0.5) + np.float32(1e-3) * (rr[:, t % 2] * np.float32(t))
                ss[:, t] = np.float32(0.5) + np.float32(1e-3) * (ss[:, t % 2] * np.float32(t))
        else:
            rr, ss = r, s
        v, _ = oracle.v0_omp(k, m, n, ss, rr)
        index = nns.DeviceIndex(dev(torch, rr))
        g = index.search(dev(torch, ss), nns.FLAG_FORCE_TENSOR | nns.FLAG_V0_ROUNDING).cpu().numpy()
        st = nns.tensor_stats()
        assert st["overflow"] == 0 and st["kp"] == want_kp and st["mode"] == want_mode, (name, st)
        assert np.array_equal(g, v), (name, int((g != v).sum()))


@pytest.mark.parametrize("k,m,n,kind", [(10, 700, 60_000, "uniform"), (13, 513, 50_000, "uniform"), (24, 1000, 70_000, "uniform"),
                                        (48, 600, 40_000, "uniform"), (64, 512, 30_000, "uniform"), (100, 300, 20_000, "uniform"),
                                        (128, 1024, 50_000, "uniform"), (126, 257, 12_345, "uniform"), (64, 400, 30_000, "scaled"),
                                        (62, 300, 9_000, "uniform"), (61, 300, 9_000, "uniform"), (125, 260, 7_000, "uniform"),
                                        (128, 300, 20_000, "scaled")])
def test_f16_mode_matches_v0(nns, oracle, torch_mod, k, m, n, kind):
    """Plain F16 operands with F16 accumulators (every 43 <= k <= 128 index built in one piece; 10 <= k <= 42 when the probe
    picks it): packed TMEM loads + HMNMX2 epilogue, scores unscaled per query before they are compared.  V0's answers
    bit for bit with V0 rounding; data far from the unit cube (coordinates ~1e4, ~1e-4) goes through the scaling.
    (k = 62..64 and 126..128 are the shapes of the NNS_T_EPI_NORM build variant: no norm columns, the epilogue adds |r'|^2.)"""
    torch = torch_mod
    s, r = make_case("uniform", k, m, n, 91)
    if kind == "scaled":
        s, r = (s * np.float32(2.5e4) - np.float32(7e3)).astype(np.float32), (r * np.float32(2.5e4) - np.float32(7e3)).astype(np.float32)
        s[::3] = (s[::3] * np.float32(4.0)).astype(np.float32)  # a third of the queries well outside the reference cloud
    v, _ = oracle.v0_omp(k, m, n, s, r)
    index = nns.DeviceIndex(dev(torch, r))
    g = index.search(dev(torch, s), nns.FLAG_FORCE_TENSOR | nns.FLAG_V0_ROUNDING).cpu().numpy()
    st = nns.tensor_stats()
    assert st["overflow"] == 0 and st["mode"] == "f16", st
    if k <= 64 and n >= 30_000:  # (small n: every reference split of a strip seeds its own candidates)
        assert st["candidates"] < m * ((n + 31) // 32) // 2, st  # the screen screens (distances concentrate as k grows: less so)
    assert np.array_equal(g, v), int((g != v).sum())
    tiny = np.float32(1e-4)
    g2 = nns.DeviceIndex(dev(torch, (r * tiny).astype(np.float32))).search(dev(torch, (s * tiny).astype(np.float32)),
                                                                          nns.FLAG_FORCE_TENSOR | nns.FLAG_V0_ROUNDING).cpu().numpy()
    v2, _ = oracle.v0_omp(k, m, n, (s * tiny).astype(np.float32), (r * tiny).astype(np.float32))
    assert np.array_equal(g2, v2) and nns.tensor_stats()["mode"] == "f16"


@pytest.mark.parametrize("k", [16, 64, 3])
@pytest.mark.parametrize("scale", [1e-21, 1e-17, 1e-12])
def test_tensor_screen_on_data_of_tiny_extent(nns, oracle, torch_mod, k, scale):
    """Coordinates of 1e-21 .. 1e-12: FP32 norms and distances are (nearly) denormal, V0's own distances quantise to a
    few values (many ties -> lowest index), the F16 scale would overflow FP32.  The screen must step aside (band = INF:
    every unit is re-scored exactly, or the FP32 kernel takes over) and the answer must still be V0's, bit for bit."""
    torch = torch_mod
    m, n = 300, 20_000
    s, r = make_case("uniform", k, m, n, 97)
    s, r = (s.astype(np.float64) * scale).astype(np.float32), (r.astype(np.float64) * scale).astype(np.float32)
    v, _ = oracle.v0_omp(k, m, n, s, r)
    g = nns.DeviceIndex(dev(torch, r)).search(dev(torch, s), nns.FLAG_FORCE_TENSOR | nns.FLAG_V0_ROUNDING).cpu().numpy()
    assert np.array_equal(g, v), (int((g != v).sum()), nns.tensor_stats())


def test_f16_mode_reference_outside_the_sampled_radius_is_not_trusted(nns, oracle, torch_mod):
    """The F16 scale comes from a strided block sample.  A reference 1000 cloud radii out that the sample missed would
    overflow the 16-bit operands: the image kernel flags the section, every query then takes the exact path (all units
    are candidates, or the FP32 kernel behind the overflow flag), and the outlier is still found by the query next to it."""
    torch = torch_mod
    k, m, n = 48, 300, 300_000  # 2344 blocks -> the sample takes every 3rd block
    s, r = make_case("uniform", k, m, n, 93)
    r[128 + 5] = np.float32(3000.0)   # block 1: not sampled
    s[7] = np.float32(2999.5)
    v, _ = oracle.v0_omp(k, m, n, s, r)
    assert v[7] == 128 + 5
    g = nns.DeviceIndex(dev(torch, r)).search(dev(torch, s), nns.FLAG_FORCE_TENSOR | nns.FLAG_V0_ROUNDING).cpu().numpy()
    assert nns.tensor_stats()["mode"] == "f16"
    assert np.array_equal(g, v), int((g != v).sum())


# ---- tcgen05 K-loop kernel, 128 < k <= 509 (tensor_longk.cu) ---------------------------------------
@pytest.mark.parametrize("k,m,n", [(129, 300, 5000), (200, 513, 12345), (256, 1024, 40000), (317, 700, 9000), (318, 700, 9000),
                                   (400, 256, 8192), (509, 1000, 6000), (192, 257, 129)])
def test_long_contractions_on_the_tensor_cores(nns, oracle, torch_mod, k, m, n):
    """k > 128: the K-loop screen (A resident in shared memory, one 64-column block of B per stage; 256 query rows
    per CTA up to k = 317, 128 above) + exact FP32 re-score must return V0's indices (V0 rounding -> identical)."""
    torch = torch_mod
    s, r = make_case("uniform", k, m, n, 81)
    v, _ = oracle.v0_omp(k, m, n, s, r)
    assert nns.plan(k, m, n, nns.FLAG_FORCE_TENSOR)["path"] == 2
    index = nns.DeviceIndex(dev(torch, r))
    g = index.search(dev(torch, s), nns.FLAG_FORCE_TENSOR | nns.FLAG_V0_ROUNDING).cpu().numpy()
    st = nns.tensor_stats()
    assert st["overflow"] == 0 and st["kp"] == 64 * ((k + 3 + 63) // 64), st
    assert np.array_equal(g, v), int((g != v).sum())
    # two reference shards (index_base) + the default FMA rounding under the tie rule
    h = (n // 2 + 127) // 128 * 128
    if 0 < h < n:
        keys = None
        for r0, r1 in ((h, n), (0, h)):
            part = nns.DeviceIndex(dev(torch, r[r0:r1]), index_base=r0)
            keys = part.new_keys(m) if keys is None else keys
            part.search_keys(dev(torch, s), keys, nns.FLAG_FORCE_TENSOR)
        g2 = nns.unpack_keys(keys, m).cpu().numpy()
        rep = oracle.check_tie_rule(k, m, n, s, r, g2, v, REL_TOL)
        assert rep["violations"] == 0 and rep["exact_match_with_v0"] >= m - 2, rep


def test_long_contraction_adversarial(nns, oracle, torch_mod):
    """k = 200 with duplicates, an offset, NaN / INF coordinates and all-identical points (overflow -> FP32 fallback)."""
    torch = torch_mod
    k, m, n = 200, 300, 4000
    s, r = make_case("uniform", k, m, n, 83)
    cases = {}
    rr = r.copy(); rr[64::64] = rr[27:27 + len(rr[64::64])]; ss = s.copy(); ss[::2] = rr[(np.arange(0, m, 2) * 37) % n]
    cases["duplicates"] = (ss, rr)
    cases["offset"] = ((s + np.float32(100.0)).astype(np.float32), (r + np.float32(100.0)).astype(np.float32))
    rn = r.copy(); rn[5, 3] = np.nan; rn[77, 0] = np.inf; rn[1000] = np.nan
    cases["nan_inf"] = (s, rn)
    cases["all_identical"] = (s, np.tile(r[:1], (n, 1)))
    for name, (a, b) in cases.items():
        v, _ = oracle.v0_omp(k, m, n, a, b)
        g = nns.DeviceIndex(dev(torch, b)).search(dev(torch, a), nns.FLAG_FORCE_TENSOR | nns.FLAG_V0_ROUNDING).cpu().numpy()
        assert np.array_equal(g, v), (name, int((g != v).sum()), nns.tensor_stats())


# ---- K nearest neighbours through the tcgen05 screen ------------------------------------------------
@pytest.mark.parametrize("kind,k,m,n,K", [("uniform", 3, 2048, 200_000, 8), ("clustered", 3, 3000, 300_000, 16), ("uniform", 16, 1024, 100_000, 32),
                                          ("uniform", 128, 512, 50_000, 5), ("grid", 3, 700, 60_000, 1), ("uniform", 200, 300, 20_000, 4)])
def test_topk_on_the_tensor_screen_matches_the_oracle(nns, oracle, torch_mod, kind, k, m, n, K):
    """nns_b200_topk_keys with the tcgen05 screen forced: exact FP32 K-nearest over a block sample fixes a distance
    threshold per query, the screen keeps the 32-reference units that can hold anything below it, their exact
    distances are merged into the sorted lists.  With V0 rounding the lists must equal the oracle's exactly --
    duplicates, grids and K = 1 (= V0) included -- and must not depend on shards or on the FP32 kernel."""
    torch = torch_mod
    s, r = make_case(kind, k, m, n, 101)
    want_i, want_d = oracle.v0_topk(k, m, n, K, s, r)
    index = nns.DeviceIndex(dev(torch, r))
    gi, gd = index.topk(dev(torch, s), K, nns.FLAG_V0_ROUNDING | nns.FLAG_FORCE_TENSOR)
    st = nns.tensor_stats()
    gi, gd = gi.cpu().numpy(), gd.cpu().numpy()
    assert np.array_equal(gi, want_i), (int((gi != want_i).sum()), st)
    assert np.array_equal(gd, want_d)
    assert st["overflow"] == 0, st
    fi, _ = index.topk(dev(torch, s), K, nns.FLAG_V0_ROUNDING | nns.FLAG_FORCE_LOWK)
    assert np.array_equal(fi.cpu().numpy(), gi)
    # offering the same references twice does not duplicate entries (idempotent merge)
    keys = index.topk_keys(dev(torch, s), K, None, nns.FLAG_V0_ROUNDING | nns.FLAG_FORCE_TENSOR)
    keys = index.topk_keys(dev(torch, s), K, keys, nns.FLAG_V0_ROUNDING | nns.FLAG_FORCE_LOWK)
    assert np.array_equal((keys.cpu().numpy().astype(np.uint64) & np.uint64(0xFFFFFFFF)).astype(np.int64), want_i)
    if K == 1:
        assert np.array_equal(gi[:, 0], oracle.v0_omp(k, m, n, s, r)[0])


def test_topk_tensor_screen_overflow_hands_over_to_the_fp32_kernel(nns, oracle, torch_mod):
    """All references identical: every unit is inside every threshold, the candidate buffer / lists overflow, and the
    FP32 K-nearest kernel launched behind the device flag must finish the search (lowest indices win the ties)."""
    torch = torch_mod
    k, m, n, K = 3, 600, 40_000, 8
    s, r = make_case("uniform", k, m, n, 103)
    r = np.tile(r[:1], (n, 1))
    want_i, want_d = oracle.v0_topk(k, m, n, K, s, r)
    gi, gd = nns.DeviceIndex(dev(torch, r)).topk(dev(torch, s), K, nns.FLAG_V0_ROUNDING | nns.FLAG_FORCE_TENSOR)
    assert nns.tensor_stats()["overflow"] == 1
    assert np.array_equal(gi.cpu().numpy(), want_i) and np.array_equal(gd.cpu().numpy(), want_d)


def test_topk_host_abi_plans_the_tensor_screen_for_large_problems(nns, oracle):
    k, m, n, K = 3, 8192, 600_000, 8  # 4.9e9 pairs
    s, r = make_case("uniform", k, m, n, 105)
    sample = np.random.default_rng(3).permutation(m)[:256]
    want_i, want_d = oracle.v0_topk(k, 256, n, K, s[sample], r)
    hi, hd = nns.search_topk_host(k, m, n, K, s, r)
    assert nns.tensor_stats()["kp"] == 16  # the screen ran
    assert (hi[sample] != want_i).mean() < 0.01  # FMA rounding may swap near-ties
    np.testing.assert_allclose(hd[sample], want_d, rtol=2e-6)
