"""the functional GPU KD-tree (csrc/kdtree.cu) that the reference's v11 was
meant to be (core.cu:1289-1451 returns zeros: its kernel body is commented out).  It is an EXACT search: on
every input the indices must be V0's -- lowest index on exact ties, NaN never wins, index 0 when nothing is
below +INF -- whatever the tree prunes."""
import os
import time

import numpy as np
import pytest

from conftest import make_case

pytestmark = pytest.mark.gpu


@pytest.fixture(params=["gpu_build", "host_build"], autouse=True)
def build_mode(request):
    """Both builds fill the same structure: Morton order on the GPU (default), median splits on the host
    (NNS_B200_TREE_HOST_BUILD=1, the reference's way); the search must return V0's answer over either."""
    old = os.environ.get("NNS_B200_TREE_HOST_BUILD")
    os.environ["NNS_B200_TREE_HOST_BUILD"] = "1" if request.param == "host_build" else "0"
    yield request.param
    if old is None:
        os.environ.pop("NNS_B200_TREE_HOST_BUILD", None)
    else:
        os.environ["NNS_B200_TREE_HOST_BUILD"] = old


@pytest.mark.parametrize("kind,k,m,n", [("uniform", 3, 1024, 65536), ("clustered", 3, 4000, 300_000), ("grid", 3, 500, 20_000),
                                        ("uniform", 1, 300, 5000), ("uniform", 2, 300, 129), ("grid", 2, 64, 100_000),
                                        ("uniform", 8, 700, 50_000), ("uniform", 16, 300, 40_000), ("uniform", 32, 100, 9000),
                                        ("uniform", 3, 50, 1), ("uniform", 3, 50, 128), ("uniform", 3, 7, 0)])
def test_tree_search_returns_v0(nns, oracle, kind, k, m, n):
    s, r = make_case(kind, k, m, max(n, 1), 91)
    r = r[:n]
    v, _ = oracle.v0_omp(k, m, n, s, r)
    tree = nns.HostTree(k, n, r)
    g, d = tree.search(m, s, return_dist=True)
    if kind == "uniform":  # default FMA rounding: the north-star tie rule; grids are exact either way
        rep = oracle.check_tie_rule(k, m, n, s, r, g, v, 1e-5)
        assert rep["violations"] == 0 and rep["exact_match_with_v0"] >= m - 1, rep
    else:
        assert np.array_equal(g, v), int((g != v).sum())
    if n > 0:
        e = s.astype(np.float64) - r[g].astype(np.float64)
        np.testing.assert_allclose(d, (e * e).sum(1), rtol=1e-5, atol=1e-12)
    # the brute-force engine gives the same indices
    assert np.array_equal(g, nns.search_host(k, m, n, s, r))
    tree.close()


def test_tree_special_values_and_duplicates(nns, oracle):
    k, m, n = 3, 600, 30_000
    s, r = make_case("clustered", k, m, n, 92)
    r = r.copy()
    r[7] = np.nan
    r[100, 1] = np.nan
    r[200, 0] = np.inf
    r[300] = -np.inf
    r[5000:5100] = r[4000]  # a run of identical points: the lowest index must win
    s = s.copy()
    s[3] = r[4000]
    s[4, 2] = np.nan       # every distance NaN -> index 0
    s[5, 0] = np.inf       # every distance +INF (or NaN against the INF references) -> index 0
    v, _ = oracle.v0_omp(k, m, n, s, r)
    tree = nns.HostTree(k, n, r)
    g = tree.search(m, s)
    assert np.array_equal(g, v), (int((g != v).sum()), np.nonzero(g != v)[0][:10])
    assert g[4] == 0 and g[5] == 0 and g[3] == min(4000, int(v[3]))
    tree.close()


def test_tree_visits_a_sliver_of_the_references_at_low_k(nns, oracle, build_mode):
    """k = 3, n = 2^22: the tree answers 65,536 queries much faster than the brute-force path of the same
    library needs for the same answer (it scans a few of the 32,768 leaves per query)."""
    k, m, n = 3, 65536, 1 << 22
    s, r = make_case("uniform", k, m, n, 93)
    t0 = time.perf_counter()
    tree = nns.HostTree(k, n, r)
    t_build = time.perf_counter() - t0
    tree.search(256, s[:256])
    t0 = time.perf_counter()
    g = tree.search(m, s)
    t_tree = time.perf_counter() - t0
    nns.search_host(k, m, n, s, r)
    t0 = time.perf_counter()
    b = nns.search_host(k, m, n, s, r)
    t_brute = time.perf_counter() - t0
    print(f"[{build_mode}] tree build {t_build * 1e3:.1f} ms, tree search {t_tree * 1e3:.2f} ms, brute force {t_brute * 1e3:.2f} ms (incl. upload of the references)")
    sample = np.random.default_rng(9).permutation(m)[:512]
    v, _ = oracle.v0_omp(k, 512, n, s[sample], r)
    rep = oracle.check_tie_rule(k, 512, n, s[sample], r, g[sample], v, 1e-5)
    assert rep["violations"] == 0 and rep["exact_match_with_v0"] >= 511, rep
    assert np.array_equal(g, b)
    assert t_tree < t_brute
    tree.close()
