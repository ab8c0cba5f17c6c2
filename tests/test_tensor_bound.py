"""CPU check of the error bound E(q) behind the tcgen05 screen (csrc/tensor_search.cu,
tensor_query_image_kernel).  The screen is exact only if, for every query q and reference r_j,

    | (S~_j + |q'|^2) - d_V0(q, r_j) |  <=  E(q)

where S~_j is what the tensor cores accumulate (BF16 hi/lo columns of the centred operands, the
three-term |r'|^2 split, FP32 accumulation) and d_V0 is V0's FP32 distance (core.cu:38-43): then the
unit that holds V0's answer is within 2E of the smallest S~ and is always re-scored.  This test
re-computes S~ in numpy with a PESSIMISTIC accumulator (sequential, truncating to FP32 after every
addition -- the hardware is at least that accurate) and checks the inequality on data that
stresses it: unit cube, clusters, large common offsets, small and mixed scales.  It mirrors the
kernel's arithmetic for E line by line; if either side changes, this test must change with it."""
import numpy as np
import pytest

from conftest import make_case

F32 = np.float32


def bf16(x):
    """round-to-nearest-even FP32 -> BF16 (kept in an FP32 container), like __float2bfloat16_rn"""
    u = np.asarray(x, dtype=F32).view(np.uint32).astype(np.uint64)
    r = ((u + 0x7FFF + ((u >> 16) & 1)) >> 16) << 16
    return r.astype(np.uint32).view(F32)


def trunc32(x64):
    """FP64 -> FP32 rounding toward zero (a truncating adder)"""
    t = x64.astype(F32)
    over = np.abs(t.astype(np.float64)) > np.abs(x64)
    return np.where(over, np.nextafter(t, F32(0)), t).astype(F32)


def geometry(k, plain=False):
    """(contraction length, data columns): split-precision columns up to k = 42 unless the index chose the
    plain BF16 layout (tensor_geom(k, plain) in csrc/tensor_search.cu; plain exists for 10 <= k <= 42)"""
    ndata = 3 * k if (k <= 42 and not plain) else k
    for kp in (16, 32, 64):
        if ndata + 3 <= kp:
            return kp, ndata
    if ndata <= 64:
        return 80, 64
    if ndata <= 128:
        return (128, ndata) if ndata + 3 <= 128 else (144, 128)
    return 64 * ((ndata + 3 + 63) // 64), ndata  # K-loop kernel: whole 64-column blocks


def screen_scores(k, s, r, plain=False):
    """(S~ [m][n], E [m], |q'|^2 exact [m]) as the kernels compute them"""
    split = k <= 42 and not plain
    kp, _ = geometry(k, plain)
    c = (r.astype(F32).sum(axis=0, dtype=F32) / F32(len(r))).astype(F32)  # any centre works; the kernel uses the FP32 mean
    qc = (s - c).astype(F32)
    rc = (r - c).astype(F32)
    # |r'|^2 and |q'|^2: FP32 FMA chains over ascending dimensions (fmaf == exact product, one rounding)
    rn = np.zeros(len(r), F32)
    for t in range(k):
        rn = (rc[:, t].astype(np.float64) * rc[:, t].astype(np.float64) + rn.astype(np.float64)).astype(F32)
    qn = np.zeros(len(s), F32)
    for t in range(k):
        qn = (qc[:, t].astype(np.float64) * qc[:, t].astype(np.float64) + qn.astype(np.float64)).astype(F32)
    a = (F32(-2.0) * qc).astype(F32)
    ah, rh = bf16(a), bf16(rc)
    al, rl = bf16((a - ah).astype(F32)), bf16((rc - rh).astype(F32))
    if split:
        A = np.concatenate([ah, ah, al], axis=1)
        B = np.concatenate([rh, rl, rh], axis=1)
    else:
        A, B = ah, rh
    n_hi = bf16(rn)
    rem = (rn - n_hi).astype(F32)
    n_mid = bf16(rem)
    n_lo = bf16((rem - n_mid).astype(F32))
    A = np.concatenate([A, np.ones((len(s), 3), F32)], axis=1)
    B = np.concatenate([B, np.stack([n_hi, n_mid, n_lo], axis=1)], axis=1)
    acc = np.zeros((len(s), len(r)), F32)
    for col in range(A.shape[1]):  # products of two BF16 numbers are exact; every addition truncates
        acc = trunc32(acc.astype(np.float64) + np.outer(A[:, col].astype(np.float64), B[:, col].astype(np.float64)))
    # E(q): tensor_query_image_kernel, same FP32 operations
    r2 = rn.max()
    aa, rmax, u24 = np.sqrt(qn).astype(F32), F32(np.sqrt(r2)), F32(5.9604645e-8)
    c_round = F32(6.2) * F32(3.8146973e-6) if split else F32(0.0078125) * F32(1.002)
    E = (c_round + F32(kp) * F32(2.04) * F32(4.7683716e-7)) * aa * rmax + F32(kp + 5) * u24 * r2 + F32(kp + 8) * u24 * (aa + rmax) * (aa + rmax)
    E = (E * F32(1.05)).astype(F32)
    qn64 = (qc.astype(np.float64) ** 2).sum(axis=1)
    return acc, E, qn64


def v0_distances(s, r):
    """V0's FP32 distances (core.cu:38-43): sub, mul, add each rounded, ascending dimensions"""
    d = np.zeros((len(s), len(r)), F32)
    for t in range(s.shape[1]):
        e = (s[:, t][:, None] - r[:, t][None, :]).astype(F32)
        d = (d + (e * e).astype(F32)).astype(F32)
    return d


CASES = ["uniform", "clustered", "offset1000", "scale1e-3", "mixed", "one_outlier"]


@pytest.mark.parametrize("k", [1, 3, 4, 9, 16, 42, 43, 64, 128, 129, 200, 320, 509, -10, -16, -29, -30, -42])
@pytest.mark.parametrize("case", CASES)
def test_screen_error_stays_inside_the_band(k, case):
    plain, k = k < 0, abs(k)  # negative = the plain BF16 layout of a k that also has the split one
    m, n = 48, 1536
    if case == "clustered" and k == 3:
        s, r = make_case("clustered", k, m, n, 11)
    else:
        s, r = make_case("uniform", k, m, n, 11)
    s, r = s.astype(np.float64), r.astype(np.float64)
    if case == "offset1000":
        s, r = s + 1000.0, r + 1000.0
    elif case == "scale1e-3":
        s, r = s * 1e-3, r * 1e-3
    elif case == "mixed":
        r[::7] *= 50.0
        s[::5] *= 0.01
    elif case == "one_outlier":
        r[3] = 300.0
    s, r = s.astype(F32), r.astype(F32)
    acc, E, qn64 = screen_scores(k, s, r, plain)
    d = v0_distances(s, r).astype(np.float64)
    err = np.abs(acc.astype(np.float64) + qn64[:, None] - d)
    worst = (err / E[:, None].astype(np.float64)).max()
    assert worst <= 1.0, f"screen error reaches {worst:.3f} x E"
    # the bound must also be worth something: on the unit cube E is a small fraction of the spread of distances
    if case == "uniform":
        assert (E.astype(np.float64) < 0.05 * d.max(axis=1)).all()


@pytest.mark.parametrize("k,case", [(3, "uniform"), (3, "clustered"), (16, "uniform"), (16, "mixed"), (128, "uniform"), (64, "offset1000"),
                                    (-16, "uniform"), (-16, "mixed"), (-30, "offset1000")])
def test_screen_and_rescore_logic_returns_v0(k, case):
    """The screen's control logic on top of the emulated scores, as the kernels run it: per query a
    running minimum over 32-reference units in index order, a unit is recorded when its minimum is
    within the band 2E of the running minimum, recorded units within 2E of the FINAL minimum are
    re-scored exactly (V0's FP32 distances), and the smallest packed (distance, index) key wins.
    Whatever the unit order and the rounding of the scores, the answer must be V0's: first minimum
    of the FP32 distances.  (Grid-snapped clustered data: many exact ties.)"""
    plain, k = k < 0, abs(k)
    m, n = 40, 2048
    if case == "clustered":
        s, r = make_case("clustered", k, m, n, 3)
    else:
        s, r = make_case("uniform", k, m, n, 3)
        if case == "offset1000":
            s, r = (s.astype(np.float64) + 1000.0).astype(F32), (r.astype(np.float64) + 1000.0).astype(F32)
        if case == "mixed":
            r = r.copy()
            r[::7] *= F32(50.0)
    acc, E, _ = screen_scores(k, s, r, plain)
    d = v0_distances(s, r)
    v0 = d.argmin(axis=1)  # numpy argmin = first minimum = V0's strict '>' update (core.cu:44)
    band = (F32(2.0) * E).astype(F32)
    total_candidates = 0
    for q in range(m):
        run_min, cand = F32(np.inf), []
        for u in range(n // 32):
            cm = acc[q, 32 * u:32 * u + 32].min()
            if cm <= run_min + band[q]:
                cand.append((u, cm))
                run_min = min(run_min, cm)
        live = [u for (u, cm) in cand if cm <= run_min + band[q]]
        total_candidates += len(live)
        best = min((d[q, j], j) for u in live for j in range(32 * u, 32 * u + 32))
        assert best[1] == v0[q], (q, best, v0[q], d[q, v0[q]])
    if case != "mixed" and not (plain and case == "offset1000"):  # references 50x farther out inflate max |r'| and with it the band: still exact, no longer selective
        assert total_candidates < m * (n // 32) // 2  # the screen screens
