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


def f16(x):
    """round-to-nearest-even -> F16 (kept in an FP32 container), like __float2half_rn; subnormals, overflow -> INF"""
    with np.errstate(over="ignore"):
        return np.asarray(x).astype(np.float16).astype(F32)


def f16_ref_scale(rmax_sampled):
    """tensor_f16_ref_scale (csrc/tensor_common.cuh): power of two that brings the sampled radius to [4, 8)"""
    if not (1e-30 < rmax_sampled < 1e30):
        return F32(1.0)
    _, e = np.frexp(F32(rmax_sampled))
    return F32(np.ldexp(1.0, 3 - int(e)))


F16_R_MAX = F32(176.0)


def f16_query_scale(A, R):
    """tensor_f16_query_scale: largest power of two t <= 2^8 with t R (R + 2A) <= 2^15 and 2 t A <= 2^15; 0 = unusable"""
    A, R = F32(A), F32(R)
    if not (A <= F32(1e30)) or not (R <= F16_R_MAX):
        return F32(0.0)
    lim = F32(256.0)
    p = F32(R * F32(R + F32(2.0) * A))
    if p > 0:
        lim = min(lim, F32(F32(32768.0) / p))
    if A > 0:
        lim = min(lim, F32(F32(16384.0) / A))
    if not (lim >= F32(6.1035156e-5)):
        return F32(0.0)
    _, e = np.frexp(F32(lim))
    return F32(np.ldexp(1.0, int(e) - 1))


EPI_NORM = False  # NNS_T_EPI_NORM of the build (csrc/tensor_common.cuh); the emulation of the variant is tested either way


def geometry(k, plain=False, en=None):
    """(contraction length, data columns): split-precision columns up to k = 42 unless the index chose the
    plain layout (tensor_geom(k, plain) in csrc/tensor_common.cuh; plain F16 exists for 10 <= k <= 128)"""
    ndata = 3 * k if (k <= 42 and not plain) else k
    en = EPI_NORM if en is None else en
    if plain and en and (61 < ndata <= 64 or 125 < ndata <= 128):  # F16 mode: no norm columns, the epilogue adds |r'|^2 (TensorGeom::en)
        return (64 if ndata <= 64 else 128), ndata
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
    c_round = F32(6.06) * F32(1.5258789e-5) if split else F32(0.015625) * F32(1.002)  # BF16 unit roundoff 2^-8
    E = (c_round + F32(kp) * F32(2.04) * F32(4.7683716e-7)) * aa * rmax + F32(kp + 5) * u24 * r2 + F32(kp + 8) * u24 * (aa + rmax) * (aa + rmax)
    E = (E * F32(1.05)).astype(F32)
    qn64 = (qc.astype(np.float64) ** 2).sum(axis=1)
    screen_scores.last = {"aa": aa, "rmax": rmax, "kp": kp}
    return acc, E, qn64


def screen_scores_f16(k, s, r, sample_step=5, epi_norm=None):
    """(S~ [m][n] unscaled, E [m], |q'|^2 exact [m], usable [m]) in the F16 mode (THDR_MODE = 2): F16 operands scaled by
    powers of two, F16 accumulator re-rounded (to nearest, as measured on B200: tools/ubench_f16acc.cu) after every
    16-column MMA step; the reference scale comes from a SAMPLE of the references (every sample_step-th), as in the
    index build, so the true radius may exceed what the scale was chosen for."""
    kp, ndata = geometry(k, True, epi_norm)
    c = (r.astype(F32).sum(axis=0, dtype=F32) / F32(len(r))).astype(F32)
    qc = (s - c).astype(F32)
    rc = (r - c).astype(F32)
    rn = np.zeros(len(r), F32)
    for t in range(k):
        rn = (rc[:, t].astype(np.float64) * rc[:, t].astype(np.float64) + rn.astype(np.float64)).astype(F32)
    qn = np.zeros(len(s), F32)
    for t in range(k):
        qn = (qc[:, t].astype(np.float64) * qc[:, t].astype(np.float64) + qn.astype(np.float64)).astype(F32)
    r2 = rn.max()
    aa, rmax = np.sqrt(qn).astype(F32), F32(np.sqrt(r2))
    sc = f16_ref_scale(F32(np.sqrt(rn[::sample_step].max())))
    tq = np.array([f16_query_scale(F32(sc * a), F32(sc * rmax)) for a in aa], F32)
    flagged = bool(((rn * sc * sc) > F16_R_MAX * F16_R_MAX).any())  # tensor_ref_image_kernel: out-of-range reference
    usable = (tq > 0) & (not flagged)
    tt = np.where(tq > 0, tq, F32(1.0)).astype(F32)
    en = ndata + 3 > kp  # no room for norm columns: one F16 norm per reference, added by an HFMA2 in the epilogue
    norm_col = kp - 16 if kp in (80, 144) else ndata  # tensor_geom: the norm columns get their own step when the data fills the blocks
    A = np.zeros((len(s), kp), F32)
    B = np.zeros((len(r), kp), F32)
    A[:, :k] = f16((F32(-2.0) * tt * sc)[:, None] * qc)
    B[:, :k] = f16(sc * rc)
    nv = (rn * sc * sc).astype(F32)
    n_hi = f16(nv)
    rem = (nv - n_hi).astype(F32)
    n_mid = f16(rem)
    n_lo = f16((rem - n_mid).astype(F32))
    if not en:
        A[:, norm_col:norm_col + 3] = f16(tt)[:, None]
        B[:, norm_col], B[:, norm_col + 1], B[:, norm_col + 2] = n_hi, n_mid, n_lo
    acc = np.zeros((len(s), len(r)), np.float64)
    with np.errstate(over="ignore", invalid="ignore"):
        for st in range(kp // 16):  # one MMA instruction: exact sum of 16 products + accumulator, rounded to F16
            blk = A[:, 16 * st:16 * st + 16].astype(np.float64) @ B[:, 16 * st:16 * st + 16].astype(np.float64).T
            acc = (acc + blk).astype(np.float16).astype(np.float64)
        if en:  # HFMA2: t * N + acc, one rounding
            acc = (acc + np.outer(f16(tt).astype(np.float64), n_hi.astype(np.float64))).astype(np.float16).astype(np.float64)
    uq = (tt * sc * sc).astype(np.float64)
    S = acc / uq[:, None]
    # E(q): tensor_error_bound_f16, same FP32 operations
    u24, u11 = F32(5.9604645e-8), F32(4.8828125e-4)
    steps = kp // 16 + (1 if en else 0)
    big = rmax * rmax + F32(2.0) * aa * rmax
    sub = (F32(np.sqrt(F32(k))) * (F32(2.0) * tt * sc * aa + sc * rmax) + F32(steps + 3)) * F32(2.9802322e-8) / (tt * sc * sc)
    E = (F32(2.0) * u11 * F32(1.001) + F32(kp) * F32(2.04) * F32(4.7683716e-7)) * aa * rmax + F32(steps) * u11 * F32(1.016) * F32(1.002) * big + sub \
        + (u11 * F32(1.001) * rmax * rmax if en else F32(0.0)) + F32(kp + 5) * u24 * r2 + F32(kp + 8) * u24 * (aa + rmax) * (aa + rmax)
    E = (E * F32(1.05)).astype(F32)
    qn64 = (qc.astype(np.float64) ** 2).sum(axis=1)
    screen_scores_f16.last = {"aa": aa, "rmax": rmax, "rmax_sampled": F32(np.sqrt(rn[::sample_step].max())), "kp": kp, "tq": tq, "sc": sc, "en": en}
    return S, E, qn64, usable


def v0_distances(s, r):
    """V0's FP32 distances (core.cu:38-43): sub, mul, add each rounded, ascending dimensions"""
    d = np.zeros((len(s), len(r)), F32)
    for t in range(s.shape[1]):
        e = (s[:, t][:, None] - r[:, t][None, :]).astype(F32)
        d = (d + (e * e).astype(F32)).astype(F32)
    return d


CASES = ["uniform", "clustered", "offset1000", "scale1e-3", "mixed", "one_outlier"]


def assert_library_bound(nns, k, mode, aa, rmax, rmax_sampled, E, kp, tq=None, sc=None):
    """the E / scales / geometry this file computes == what the shipped library computes for the same inputs
    (nns_b200_tensor_bound, a host-side pure function): the emulation tests the product's formula, not a copy of it"""
    for i in range(0, len(aa), max(1, len(aa) // 8)):
        b = nns.tensor_bound(k, mode, float(aa[i]), float(rmax), float(rmax_sampled))
        assert b["kp"] == kp, (b, kp)
        if mode == 2:
            assert b["s"] == float(sc) and b["t"] == float(tq[i]), (b, float(sc), float(tq[i]))
            if tq[i] == 0:
                continue
        assert abs(b["E"] - float(E[i])) <= 4e-6 * float(E[i]), (b, float(E[i]))  # same FP32 formula; association may differ by an ulp or two


@pytest.mark.parametrize("k", [1, 3, 4, 9, 16, 42, 43, 64, 128, 129, 200, 320, 509, -10, -13, -16, -29, -30, -42, -61, -62, -64, -100, -125, -126, -128])
@pytest.mark.parametrize("case", CASES)
def test_screen_error_stays_inside_the_band(nns, k, case):
    plain, k = k < 0, abs(k)  # negative = the plain F16 mode (10 <= k <= 128)
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
    d = v0_distances(s, r).astype(np.float64)
    if plain:
        acc, E, qn64, usable = screen_scores_f16(k, s, r)
        if case == "one_outlier":  # the sample misses the outlier, the scale is 300x off: the section must be flagged, not trusted
            assert not usable.any()
            return
        assert usable.all()
        assert np.isfinite(acc).all()  # no operand, no partial sum overflowed
        L = screen_scores_f16.last
        assert_library_bound(nns, k, 2, L["aa"], L["rmax"], L["rmax_sampled"], E, L["kp"], L["tq"], L["sc"])
    else:
        acc, E, qn64 = screen_scores(k, s, r, plain)
        L = screen_scores.last
        assert_library_bound(nns, k, 0, L["aa"], L["rmax"], L["rmax"], E, L["kp"])
    err = np.abs(acc.astype(np.float64) + qn64[:, None] - d)
    worst = (err / E[:, None].astype(np.float64)).max()
    assert worst <= 1.0, f"screen error reaches {worst:.3f} x E"
    # the bound must also be worth something: on the unit cube E is a small fraction of the spread of distances
    if case == "uniform":
        assert (E.astype(np.float64) < 0.05 * d.max(axis=1)).all()


@pytest.mark.parametrize("k", [62, 64, 126, 128])
@pytest.mark.parametrize("case", ["uniform", "offset1000", "scale1e-3", "mixed"])
def test_epilogue_norm_variant_stays_inside_the_band(k, case):
    """NNS_T_EPI_NORM = 1: no norm columns, one F16 norm per reference added by an HFMA2 in the epilogue"""
    s, r = make_case("uniform", k, 32, 1024, 13)
    s, r = s.astype(np.float64), r.astype(np.float64)
    if case == "offset1000":
        s, r = s + 1000.0, r + 1000.0
    elif case == "scale1e-3":
        s, r = s * 1e-3, r * 1e-3
    elif case == "mixed":
        r[::7] *= 50.0
        s[::5] *= 0.01
    s, r = s.astype(F32), r.astype(F32)
    acc, E, qn64, usable = screen_scores_f16(k, s, r, epi_norm=True)
    assert usable.all() and np.isfinite(acc).all()
    d = v0_distances(s, r).astype(np.float64)
    worst = (np.abs(acc + qn64[:, None] - d) / E[:, None].astype(np.float64)).max()
    assert worst <= 1.0, f"screen error reaches {worst:.3f} x E"


@pytest.mark.parametrize("k,case", [(3, "uniform"), (3, "clustered"), (16, "uniform"), (16, "mixed"), (128, "uniform"), (64, "offset1000"),
                                    (-16, "uniform"), (-16, "mixed"), (-30, "offset1000"), (-128, "uniform"), (-64, "clustered")])
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
    if plain:
        acc, E, _, usable = screen_scores_f16(k, s, r)
        assert usable.all()
        acc = acc.astype(F32)  # the epilogue unscales in FP32 (exact: powers of two)
    else:
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
