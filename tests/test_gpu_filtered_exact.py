"""The filtered exact predicate of the scoring kernels (csrc/score_h_filt.cuh, csrc/score_p_filt.cuh).

B2R_ARITH_EXACT takes the sign of a division-free FMA margin wherever a proved error bound separates it from OpenCV's own
comparison and runs OpenCV's un-fused sequence elsewhere.  It must return the SAME integers as that sequence run on every
evaluation (B2R_ARITH_EXACT_UNFILTERED) and as the CPU oracle (HomographyEstimatorCallback::computeError / findInliers,
SURVEY.md A.5; reference call site main_v1.py:312) — also when the threshold is put exactly on an evaluation's error, one
float below it, one above it, when hypotheses are degenerate, huge, tiny, NaN or infinite, when a
denominator changes sign inside the point range, and when points are not finite.  All comparisons are equalities."""
import numpy as np
import pytest

import ransac_b200
from ransac_b200 import synth

pytestmark = pytest.mark.gpu


def _problem(n, outliers, seed, noise=1.0):
    rng = np.random.default_rng(seed)
    s, d, _ = synth.homography_set(n, outliers, rng, noise_px=noise)
    return np.asarray(s, dtype=np.float32), np.asarray(d, dtype=np.float32)


def _models_from_samples(ctx, sq, dq, n_models, seed):
    rng = np.random.default_rng(seed)
    idx = np.stack([rng.choice(len(sq), 4, replace=False) for _ in range(n_models)]).astype(np.int32)
    H, ok, _ = ctx.solve_h4(sq, dq, idx)
    m = H.reshape(-1, 9)[:, :8].astype(np.float32)
    m[~ok] = np.nan
    return m


def _errors_f32(models, src, dst):
    """OpenCV's un-fused fp32 sequence in numpy (every operation rounded to float32), [H, N]."""
    m = models.astype(np.float32)[:, None, :]
    X, Y = src[None, :, 0], src[None, :, 1]
    u, v = dst[None, :, 0], dst[None, :, 1]
    one = np.float32(1.0)
    with np.errstate(all="ignore"):
        ww = one / ((m[..., 6] * X + m[..., 7] * Y) + one)
        dx = ((m[..., 0] * X + m[..., 1] * Y) + m[..., 2]) * ww - u
        dy = ((m[..., 3] * X + m[..., 4] * Y) + m[..., 5]) * ww - v
        return dx * dx + dy * dy


def _check(ctx, oracle, models, src, dst, thr_sq):
    with np.errstate(all="ignore"):
        ref = oracle.h_count_inliers_f32(models, src, dst, thr_sq)
    got = ctx.score_h(models, src, dst, thr_sq, ransac_b200.ARITH_EXACT)
    unf = ctx.score_h(models, src, dst, thr_sq, ransac_b200.ARITH_EXACT_UNFILTERED)
    np.testing.assert_array_equal(unf, ref)
    np.testing.assert_array_equal(got, ref)
    return ref


def test_threshold_on_below_and_above_an_evaluation(ctx, oracle):
    """thr = the fp32 error of one (hypothesis, point), the float below it and the float above it: the count of that
    hypothesis changes by exactly that point between the three, and every count equals the oracle's."""
    sq, dq = _problem(3000, 0.5, 301)
    models = _models_from_samples(ctx, sq, dq, 1200, 302)
    e = _errors_f32(models, sq, dq)
    rng = np.random.default_rng(303)
    finite = np.isfinite(e)
    near = np.argwhere(finite & (e > 0.5) & (e < 200.0))
    picks = near[rng.choice(len(near), 24, replace=False)]
    for (k, p) in picks:
        t = np.float32(e[k, p])
        counts = []
        for thr_sq in (np.nextafter(t, np.float32(0)), t, np.nextafter(t, np.float32(np.inf))):
            counts.append(_check(ctx, oracle, models, sq, dq, np.float32(thr_sq))[k])
        assert counts[1] >= counts[0] + 1 and counts[2] >= counts[1]   # the point itself enters at thr = its error


def test_equal_on_mixed_hypotheses(ctx, oracle):
    """Near-truth hypotheses (dense near the threshold), sample hypotheses, and pathological rows: NaN, one NaN
    coefficient, infinities, 1e20 / 1e-20 scales, all zeros, denominators that vanish inside the point range."""
    n = 5000 + 37   # full tiles + a ragged tail
    sq, dq = _problem(n, 0.4, 311)
    models = _models_from_samples(ctx, sq, dq, 3000, 312)
    good = np.argwhere(~np.isnan(models).any(axis=1)).ravel()
    e = _errors_f32(models[good], sq, dq)
    best = good[np.argsort((e <= 9.0).sum(axis=1))[-40:]]
    rng = np.random.default_rng(313)
    near = np.concatenate([models[best] * (1 + 1e-4 * rng.standard_normal((40, 8)).astype(np.float32)) for _ in range(20)])
    odd = np.tile(models[best[0]], (16, 1)).astype(np.float32)
    odd[0] = np.nan
    odd[1, 3] = np.nan
    odd[2, 0] = np.inf
    odd[3, 7] = -np.inf
    odd[4] *= np.float32(1e20)
    odd[5, :6] *= np.float32(1e-20)
    odd[6] = 0
    odd[7, 6:8] = (np.float32(-1.0 / sq[:, 0].mean()), 0)        # w = 0 near the middle of the X range
    odd[8, 6:8] = (0, np.float32(-1.0 / sq[:, 1].mean()))
    odd[9, 6:8] = (np.float32(3e38), np.float32(3e38))
    odd[10, :6] = 0
    odd[11, 2] = np.float32(2.0 ** 41)
    odd[12, 6] = np.float32(2.0 ** 41)
    odd[13, 0] = np.float32(1e-42)                                  # a denormal coefficient
    allm = np.concatenate([models, near.astype(np.float32), odd])
    for thr in (0.7, 3.0, 75.0):
        _check(ctx, oracle, allm, sq, dq, np.float32(thr * thr))


@pytest.mark.parametrize("thr_sq", [0.0, 1e-45, 2.0 ** -41, 2.0 ** -39, 2.0 ** 39, 2.0 ** 41, 3e38, np.inf])
def test_threshold_range(ctx, oracle, thr_sq):
    """Thresholds inside and outside the guards of the bound (2^-40 .. 2^40): the counts do not depend on the route."""
    sq, dq = _problem(2100, 0.5, 321)
    models = _models_from_samples(ctx, sq, dq, 600, 322)
    _check(ctx, oracle, models, sq, dq, np.float32(thr_sq))


def test_non_finite_and_huge_points(ctx, oracle):
    """A NaN / infinite / 2^41 coordinate in a tile sends that tile through the un-fused sequence; the other tiles
    stay filtered; counts are the oracle's."""
    sq, dq = _problem(3100, 0.5, 331)
    models = _models_from_samples(ctx, sq, dq, 900, 332)
    for where, value in ((5, np.nan), (1500, np.inf), (3099, -np.inf), (2047, np.float32(2.0 ** 41))):
        for arr in (0, 1):
            s2, d2 = sq.copy(), dq.copy()
            (s2 if arr == 0 else d2)[where, arr] = value
            _check(ctx, oracle, models, s2, d2, np.float32(9.0))


def test_full_size_equals_unfiltered(ctx):
    """BASELINE configs[2] shape (100k x 100k): every one of the 100 000 counts equal between the two routes."""
    src, dst = synth.config_homography(2)
    sq, dq = np.asarray(src, dtype=np.float32), np.asarray(dst, dtype=np.float32)
    models = _models_from_samples(ctx, sq, dq, 100_000, 342)
    thr_sq = np.float32(9.0)
    a = ctx.score_h(models, sq, dq, thr_sq, ransac_b200.ARITH_EXACT)
    b = ctx.score_h(models, sq, dq, thr_sq, ransac_b200.ARITH_EXACT_UNFILTERED)
    np.testing.assert_array_equal(a, b)
    assert a.max() > 40_000


# ---- 3x4 models (cv2.solvePnPRansac, main_v1.py:497-502): csrc/score_p_filt.cuh ---------------------------------------------

K = synth.K_1898


def _pnp_errors(models, P32, px32, K):
    """cv::projectPoints (fp64, un-fused) + the fp32 error, as csrc/score_p.cuh::p_inlier_exact runs them, [H, N]."""
    fx, fy, cx, cy = K[0, 0], K[1, 1], K[0, 2], K[1, 2]
    X, Y, Z = (P32[:, i].astype(np.float64)[None, :] for i in range(3))
    m = models[:, None, :]
    with np.errstate(all="ignore"):
        x = ((m[..., 0] * X + m[..., 1] * Y) + m[..., 2] * Z) + m[..., 9]
        y = ((m[..., 3] * X + m[..., 4] * Y) + m[..., 5] * Z) + m[..., 10]
        z = ((m[..., 6] * X + m[..., 7] * Y) + m[..., 8] * Z) + m[..., 11]
        z = np.where(z != 0, 1.0 / z, 1.0)
        pu = ((x * z) * fx + cx).astype(np.float32)
        pv = ((y * z) * fy + cy).astype(np.float32)
        dx, dy = px32[None, :, 0] - pu, px32[None, :, 1] - pv
        return dx * dx + dy * dy


def _pnp_check(ctx, models, P, px, thr_sq, ref=None):
    got = ctx.score_p(models, P, px, K, thr_sq, ransac_b200.ARITH_EXACT)
    unf = ctx.score_p(models, P, px, K, thr_sq, ransac_b200.ARITH_EXACT_UNFILTERED)
    np.testing.assert_array_equal(got, unf)
    if ref is not None:
        np.testing.assert_array_equal(unf, ref)
    return unf


def _poses(oracle, rng, n_near, n_wild):
    """Poses around the true one (the rotation is perturbed about the CAMERA, t = -R cam: errors of a few pixels, dense
    near the threshold) and wild ones (cameras inside and around the cloud: the depth changes sign among the points)."""
    R0, _ = synth.look_at_pose()
    rv0 = oracle.rodrigues_inv(R0)
    out = []
    for k in range(n_near):
        R = oracle.rodrigues(rv0 + rng.normal(0, 2e-4 if k % 2 else 2e-3, 3))
        cam = synth.CAMERA_ORIGIN + rng.normal(0, 0.3 if k % 2 else 3.0, 3)
        out.append(np.concatenate([R.ravel(), -R @ cam]))
    centre = 0.5 * (synth.BOX_LO + synth.BOX_HI)
    for k in range(n_wild):
        R = oracle.rodrigues(rng.normal(0, 2.0, 3))
        cam = centre + rng.normal(0, 150.0 if k % 2 else 600.0, 3)
        out.append(np.concatenate([R.ravel(), -R @ cam]))
    return np.array(out)


def test_pnp_equal_on_mixed_hypotheses(ctx, oracle):
    rng = np.random.default_rng(401)
    n = 5000 + 37
    P, px, _ = synth.pnp_set(n, 0.4, rng)
    models = _poses(oracle, rng, 900, 1200)
    odd = np.tile(models[0], (10, 1))
    odd[0] = np.nan
    odd[1, 4] = np.nan
    odd[2, 9] = np.inf
    odd[3, 11] = 1e20
    odd[4, :9] *= 1e20
    odd[5, 9:] = 0
    odd[6, :9] = 0
    odd[7, 6:9] = 0; odd[7, 11] = 0          # z == 0 for every point: OpenCV takes z = 1
    odd[8, 6:9] = 0; odd[8, 11] = 1e-300     # 1/z overflows
    odd[9, 11] = -odd[9, 11]
    allm = np.concatenate([models, odd])
    P32, px32 = P.astype(np.float32), px.astype(np.float32)
    e = _pnp_errors(allm, P32, px32, K)
    for thr in (1.0, 8.0, 30.0):
        thr_sq = np.float32(thr * thr)
        with np.errstate(all="ignore"):
            ref = (e <= thr_sq).sum(axis=1)
        _pnp_check(ctx, allm, P, px, thr_sq, ref)
    k = 17   # spot check of the numpy restatement against the C oracle
    assert oracle.pnp_count_inliers(allm[k, :9], allm[k, 9:], K, P32, px32, 8.0)[0] == (e[k] <= np.float32(64.0)).sum()


def test_pnp_threshold_on_below_and_above_an_evaluation(ctx, oracle):
    rng = np.random.default_rng(411)
    P, px, _ = synth.pnp_set(3000, 0.4, rng)
    models = _poses(oracle, rng, 500, 300)
    P32, px32 = P.astype(np.float32), px.astype(np.float32)
    e = _pnp_errors(models, P32, px32, K)
    near = np.argwhere(np.isfinite(e) & (e > 0.5) & (e < 2000.0))
    picks = near[rng.choice(len(near), 16, replace=False)]
    for (k, p) in picks:
        t = np.float32(e[k, p])
        counts = []
        for thr_sq in (np.nextafter(t, np.float32(0)), t, np.nextafter(t, np.float32(np.inf))):
            with np.errstate(all="ignore"):
                ref = (e <= np.float32(thr_sq)).sum(axis=1)
            counts.append(_pnp_check(ctx, models, P, px, np.float32(thr_sq), ref)[k])
        assert counts[1] >= counts[0] + 1 and counts[2] >= counts[1]


def test_pnp_depth_exactly_zero(ctx, oracle):
    """Planar object points (Z = 0) under R = I, t = (tx, ty, 0): the fp64 depth is exactly 0 for every point and
    OpenCV's `z != 0 ? 1/z : 1` takes z = 1, so pixels u = fx (X + tx) + cx ARE inliers.  The margin's form assumes the
    first branch: the depth guard must send these evaluations to OpenCV's sequence."""
    rng = np.random.default_rng(421)
    n = 1500
    P = np.zeros((n, 3))
    P[:, :2] = rng.uniform(-0.3, 0.3, (n, 2))
    tx, ty = 0.05, -0.02
    px = np.stack([K[0, 0] * (P[:, 0].astype(np.float32) + tx) + K[0, 2], K[1, 1] * (P[:, 1].astype(np.float32) + ty) + K[1, 2]], axis=1)
    px[::3] += rng.normal(0, 40.0, (len(px[::3]), 2))
    models = np.zeros((6, 12))
    for k in range(6):
        models[k, :9] = np.eye(3).ravel()
        models[k, 9:] = (tx + 1e-4 * k, ty, 0.0)
    models[4, 11] = 1e-7      # a tiny non-zero depth
    models[5, 11] = 1.0
    P32, px32 = P.astype(np.float32), px.astype(np.float32)
    e = _pnp_errors(models, P32, px32, K)
    for thr_sq in (np.float32(4.0), np.float32(64.0)):
        ref = (e <= thr_sq).sum(axis=1)
        assert ref[0] > n // 2
        _pnp_check(ctx, models, P, px, thr_sq, ref)


@pytest.mark.parametrize("thr_sq", [0.0, 2.0 ** -41, 2.0 ** 41, np.inf])
def test_pnp_threshold_range(ctx, oracle, thr_sq):
    rng = np.random.default_rng(431)
    P, px, _ = synth.pnp_set(2100, 0.4, rng)
    models = _poses(oracle, rng, 200, 200)
    _pnp_check(ctx, models, P, px, np.float32(thr_sq))


def test_pnp_full_size_equals_unfiltered(ctx, oracle):
    """100k points x 100k poses: every count equal between the two routes."""
    rng = np.random.default_rng(441)
    P, px, _ = synth.pnp_set(100_000, 0.5, rng)
    models = _poses(oracle, rng, 2000, 3000)
    models = np.tile(models, (20, 1))
    models[:, 9:] += rng.normal(0, 0.05, (len(models), 3))
    models[7::1000] = np.nan
    a = ctx.score_p(models, P, px, K, np.float32(64.0), ransac_b200.ARITH_EXACT)
    b = ctx.score_p(models, P, px, K, np.float32(64.0), ransac_b200.ARITH_EXACT_UNFILTERED)
    np.testing.assert_array_equal(a, b)
    assert a.max() > 40_000


def test_pnp_non_finite_points_and_huge_intrinsics(ctx, oracle):
    """A NaN / infinite object point or pixel sends its tile through OpenCV's sequence (the other tiles stay filtered); an
    intrinsics matrix beyond the guards (2^41) sends every tile there: the two routes return the same counts."""
    rng = np.random.default_rng(451)
    P, px, _ = synth.pnp_set(3100, 0.4, rng)
    models = _poses(oracle, rng, 150, 150)
    for where, value in ((5, np.nan), (1500, np.inf), (3099, -np.inf)):
        P2, px2 = P.copy(), px.copy()
        P2[where, 1] = value
        _pnp_check(ctx, models, P2, px, np.float32(64.0))
        px2[where, 0] = value
        _pnp_check(ctx, models, P, px2, np.float32(64.0))
    Kbig = K.copy()
    Kbig[0, 0] = 2.0 ** 41
    got = ctx.score_p(models, P, px, Kbig, np.float32(64.0), ransac_b200.ARITH_EXACT)
    unf = ctx.score_p(models, P, px, Kbig, np.float32(64.0), ransac_b200.ARITH_EXACT_UNFILTERED)
    np.testing.assert_array_equal(got, unf)
