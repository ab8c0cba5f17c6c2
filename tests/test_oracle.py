"""CPU tests (-m "not gpu"): the oracle restatement against the reference's own V0 (oracle/_ref,
when built), the committed golden fixtures produced by that reference code, the known-answer
tests authored from V0's source, and the FP64 tie-rule checker."""
import os

import numpy as np
import pytest

from conftest import make_case

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "v0_golden.npz")


def golden_cases():
    z = np.load(GOLDEN)
    ci = 0
    while f"case{ci}_meta" in z:
        k, m, n, seed = (int(x) for x in z[f"case{ci}_meta"])
        yield str(z[f"case{ci}_kind"]), k, m, n, seed, z[f"case{ci}_idx"], z[f"case{ci}_sum"]
        ci += 1


def test_known_answers_from_v0_source(oracle):
    # core.cu:44 strict '>' => lowest index among exact ties; NaN never wins; n == 0 -> 0
    r = np.array([[5, 5, 5], [1, 1, 1], [1, 1, 1], [1, 1, 1], [9, 9, 9]], np.float32)
    q = np.array([[1, 1, 1], [0, 0, 0], [np.nan, 0, 0], [np.inf, 0, 0], [9, 9, 8]], np.float32)
    assert oracle.v0(3, 5, 5, q, r).tolist() == [1, 1, 0, 0, 4]
    assert oracle.v0(3, 5, 0, q, r[:0]).tolist() == [0, 0, 0, 0, 0]
    # a NaN reference point is skipped, not selected
    r2 = np.array([[np.nan, 0, 0], [2, 2, 2], [0.5, 0.5, 0.5]], np.float32)
    assert oracle.v0(3, 2, 3, q[:2], r2).tolist() == [2, 2]


def test_restatement_matches_reference_v0_bit_for_bit(oracle):
    if oracle.ref() is None:
        pytest.skip("oracle/_ref not built (no /root/reference here); golden fixtures cover this")
    for kind, k, m, n, seed in [("uniform", 3, 300, 5000, 5), ("grid", 3, 200, 3000, 6), ("uniform", 16, 64, 4096, 7),
                                ("uniform", 128, 16, 2000, 8), ("clustered", 3, 256, 4096, 9), ("uniform", 7, 1, 1, 10)]:
        s, r = make_case(kind, k, m, n, seed)
        assert np.array_equal(oracle.v0(k, m, n, s, r), oracle.ref_v0(k, m, n, s, r)), (kind, k, m, n)
        a, _ = oracle.v0_omp(k, m, n, s, r)
        b, _ = oracle.ref_v0_omp(k, m, n, s, r)
        assert np.array_equal(a, b)


def test_restatement_matches_golden_vectors(oracle):
    ncases = 0
    for kind, k, m, n, seed, idx, sums in golden_cases():
        s, r = make_case(kind, k, m, n, seed)
        assert np.allclose([s.astype(np.float64).sum(), r.astype(np.float64).sum()], sums, rtol=0, atol=0), "generator drifted"
        assert np.array_equal(oracle.v0(k, m, n, s, r), idx), (kind, k, m, n)
        ncases += 1
    assert ncases >= 20


def test_reference_rand_generator_golden(oracle, datagen):
    z = np.load(GOLDEN)
    for si in range(2):
        k, m, n = (int(x) for x in z[f"rand{si}_meta"])
        s, r = datagen.reference_rand_sample(k, m, n, 1000)
        assert np.array_equal(oracle.v0(k, m, n, s, r), z[f"rand{si}_idx"])


def test_openmp_wrapper_is_identical_to_serial(oracle):
    s, r = make_case("grid", 3, 513, 4000, 21)
    a = oracle.v0(3, 513, 4000, s, r)
    for chunk in (1, 7, 64, 1000):
        b, threads = oracle.v0_omp(3, 513, 4000, s, r, chunk)
        assert threads >= 1 and np.array_equal(a, b)


def test_tie_rule_checker(oracle):
    k, m, n = 3, 64, 2000
    s, r = make_case("uniform", k, m, n, 33)
    v = oracle.v0(k, m, n, s, r)
    rep = oracle.check_tie_rule(k, m, n, s, r, v, v)
    assert rep["violations"] == 0 and rep["exact_match_with_v0"] == m and rep["oracle_anomalies"] == 0
    bad = v.copy()
    bad[3] = (bad[3] + 1) % n
    rep = oracle.check_tie_rule(k, m, n, s, r, bad, v)
    assert rep["violations"] == 1 and rep["outside_band"] == 1
    oob = v.copy()
    oob[0] = n
    assert oracle.check_tie_rule(k, m, n, s, r, oob, v)["out_of_range"] == 1
    # exact ties: a higher index with the same FP64 distance is a violation of rule (3)
    r2 = np.vstack([r, r[v[5]:v[5] + 1]])  # duplicate of query 5's winner appended at index n
    g = oracle.v0(k, m, n + 1, s, r2)
    assert g[5] == v[5]
    hi = g.copy()
    hi[5] = n
    rep = oracle.check_tie_rule(k, m, n + 1, s, r2, hi, g)
    assert rep["higher_index_on_exact_tie"] == 1 and rep["violations"] == 1
    # near tie inside the 1e-5 band (but not exact) is accepted
    r3 = np.vstack([r, (r[v[7]].astype(np.float64) + (s[7] - r[v[7]]) * 1e-7).astype(np.float32)[None, :]])
    alt = oracle.v0(k, m, n + 1, s, r3).copy()
    other = n if alt[7] == v[7] else v[7]
    alt2 = alt.copy()
    alt2[7] = other
    rep = oracle.check_tie_rule(k, m, n + 1, s, r3, alt2, alt)
    assert rep["outside_band"] == 0


def test_clustered_workload_has_exact_ties(oracle, datagen):
    q, r = datagen.clustered_workload(512, 4096, 3, 1000)
    assert np.array_equal(r[64], r[27]) and np.array_equal(r[128], r[91])
    # every 2nd query is a copy of a reference point: its distance is exactly 0 and V0 returns
    # the lowest index among the coincident points
    v = oracle.v0(3, 512, 4096, q, r)
    for i in range(0, 512, 2):
        j = v[i]
        assert np.array_equal(r[j], q[i])
        same = np.where((r == q[i]).all(axis=1))[0]
        assert j == same.min()


def test_topk_oracle_first_neighbour_is_v0_and_lists_are_sorted(oracle):
    """oracle_v0_topk (extension): K = 1 equals V0; lists are ordered by (distance, index) and agree with
    a numpy restatement of V0's arithmetic."""
    from conftest import make_case

    for kind, k, m, n, K in [("grid", 3, 40, 500, 8), ("uniform", 16, 25, 300, 32), ("clustered", 3, 64, 129, 5), ("uniform", 5, 7, 3, 4)]:
        s, r = make_case(kind, k, m, n, 13)
        idx, dist = oracle.v0_topk(k, m, n, K, s, r)
        assert np.array_equal(idx[:, 0], oracle.v0(k, m, n, s, r))
        for i in range(m):
            acc = np.zeros(n, dtype=np.float32)
            for t in range(k):
                d = (s[i, t] - r[:, t]).astype(np.float32)
                acc = (acc + (d * d).astype(np.float32)).astype(np.float32)
            order = np.lexsort((np.arange(n), acc))[:K]
            kk = min(K, n)
            assert np.array_equal(idx[i, :kk], order[:kk])
            assert np.array_equal(dist[i, :kk], acc[order[:kk]])
            assert np.all(idx[i, kk:] == -1) and np.all(np.isinf(dist[i, kk:]))
