"""GPU: BASELINE.json's full-size configurations through size-independent properties (the oracle would need minutes
to hours there): configs[2] = 100k points x 100k hypotheses, configs[3] = 1M points x 1M hypotheses sharded 8 ways,
configs[4] = 4096 problems x 2k points x 2000 hypotheses.

Properties: (a) the returned mask is exactly the reference's fp32 un-fused error formula applied to the
returned H — recomputed here with NumPy float32 arithmetic, which rounds every operation like the reference's scalar
code; (b) the winner of a hypothesis-sharded run equals the winner of the unsharded run (counter-based sampler);
(c) the best count is the maximum of the per-hypothesis counts and the lowest id attaining it; (d) fast arithmetic
finds the same consensus set up to threshold-borderline points; and (e) a direct ORACLE comparison that stays cheap at
any size: the per-hypothesis inlier counts of 64 random hypotheses + the winner, read back from the device after the
full-size launch (b2r_h_problem_peek_hyps), equal the CPU oracle's counts of those fp32 models over all points."""
import numpy as np
import pytest

import ransac_b200
from ransac_b200 import synth

pytestmark = pytest.mark.gpu


def reference_mask_f32(H, src, dst, thr):
    """OpenCV's computeError + findInliers in float32, operation for operation (no FMA in NumPy's scalar kernels)."""
    Hf = np.asarray(H, dtype=np.float64).reshape(9).astype(np.float32)
    X, Y = src[:, 0].astype(np.float32), src[:, 1].astype(np.float32)
    u, v = dst[:, 0].astype(np.float32), dst[:, 1].astype(np.float32)
    one = np.float32(1.0)
    with np.errstate(all="ignore"):
        ww = one / ((Hf[6] * X + Hf[7] * Y) + one)
        dx = ((Hf[0] * X + Hf[1] * Y) + Hf[2]) * ww - u
        dy = ((Hf[3] * X + Hf[4] * Y) + Hf[5]) * ww - v
        err = dx * dx + dy * dy
    return (err <= np.float32(thr * thr)).astype(np.uint8)


@pytest.mark.parametrize("cfg,shards", [(2, 4), (3, 8)])
def test_large_single_problem_properties(ctx, oracle, cfg, shards):
    c = synth.CONFIGS[cfg]
    src, dst = synth.config_homography(cfg)
    n, Htot, thr = c["n_points"], c["hypotheses"], 3.0
    kw = dict(sampler=ransac_b200.SAMPLER_PHILOX, seed=1898 + cfg, solver=ransac_b200.SOLVER_FAST)
    prob = ctx.upload(src, dst)
    # unsharded, exact arithmetic
    p = ransac_b200.make_params(thr, Htot, arith=ransac_b200.ARITH_EXACT, **kw)
    prob.run(p)
    H0, m0, i0 = prob.fetch()
    assert i0[0]["status"] == 0 and i0[0]["best_count"] > 0.9 * (1 - c["outliers"]) * n * 0.9
    np.testing.assert_array_equal(m0[0], reference_mask_f32(H0[0], src, dst, thr))                      # (a)
    # (e) ORACLE comparison at full size: the counts K3 left on the device for 64 random hypotheses + the winner, against
    # the CPU oracle's count of the very fp32 models the kernel scored (exact arithmetic: equality)
    ids = np.unique(np.r_[np.random.default_rng(cfg).integers(0, Htot, 64), i0[0]["best_iter"]])
    peek = [prob.peek_hyps(int(g), 1) for g in ids]
    models0 = np.concatenate([p[1] for p in peek])
    counts0 = np.concatenate([p[2] for p in peek])
    valid = ~np.isnan(models0).any(axis=1)
    assert valid.sum() >= 48
    want = oracle.h_count_inliers_f32(models0[valid], src, dst, np.float32(thr * thr))
    np.testing.assert_array_equal(counts0[valid], want)
    assert (counts0[~valid] == 0).all()
    assert counts0[list(ids).index(i0[0]["best_iter"])] == i0[0]["best_count"]
    # sharded by hypothesis id (what 8 ranks do), MAX of the packed keys
    per = Htot // shards
    keys = [prob.score_shard(ransac_b200.make_params(thr, per, arith=ransac_b200.ARITH_EXACT, hyp_begin=r * per, **kw)) for r in range(shards)]
    best = np.maximum.reduce(keys)
    assert int(best[0]) >> 32 == i0[0]["best_count"]                                                  # (c)
    assert 0xFFFFFFFF - (int(best[0]) & 0xFFFFFFFF) == i0[0]["best_iter"]
    prob.finish(ransac_b200.make_params(thr, per, arith=ransac_b200.ARITH_EXACT, hyp_begin=0, **kw), best)
    H1, m1, i1 = prob.fetch()
    np.testing.assert_array_equal(H1, H0)                                                               # (b)
    np.testing.assert_array_equal(m1, m0)
    assert i1[0]["sample"] == i0[0]["sample"]
    # fast arithmetic
    prob.run(ransac_b200.make_params(thr, Htot, arith=ransac_b200.ARITH_FAST, **kw))
    Hf, mf, i_f = prob.fetch()
    assert abs(i_f[0]["best_count"] - i0[0]["best_count"]) <= max(3, n // 20000)                       # (d)
    # fast arithmetic against the oracle on the same 65 hypotheses: FMA margins may flip threshold-borderline points only
    peek = [prob.peek_hyps(int(g), 1) for g in ids]
    models_f = np.concatenate([p[1] for p in peek])
    counts_f = np.concatenate([p[2] for p in peek])
    np.testing.assert_array_equal(models_f[valid], models0[valid])       # the same hypotheses (sampler + solver are shared)
    assert np.abs(counts_f[valid].astype(np.int64) - want).max() <= max(3, n // 20000)
    if i_f[0]["sample"] == i0[0]["sample"]:
        assert np.abs(Hf - H0).max() / np.abs(H0).max() < 1e-5
    np.testing.assert_array_equal(mf[0], reference_mask_f32(Hf[0], src, dst, thr))                      # the final mask is always exact
    prob.free()


def test_multi_query_batch_properties(ctx):
    """configs[4]: 4096 independent problems x 2k points, 2000 hypotheses each, one batched call."""
    c = synth.CONFIGS[4]
    Q, n, thr = 4096, c["n_points"], 3.0
    rng = np.random.default_rng(1898 + 4)
    s1, d1, _ = synth.homography_set(n, c["outliers"], rng)
    src = np.broadcast_to(s1, (Q, n, 2)).copy()
    dst = d1[None] + rng.normal(0, 0.3, (Q, n, 2))
    H, ok, mask, infos = ctx.find_homography_batch(src, dst, thr, max_iters=c["hypotheses"], sampler=ransac_b200.SAMPLER_PHILOX, seed=4)
    assert ok.all()
    for q in range(0, Q, 97):
        np.testing.assert_array_equal(mask[q], reference_mask_f32(H[q], src[q], dst[q], thr))
        assert infos[q]["best_count"] > 0.5 * n
    # a problem solved alone gives the same answer as inside the batch (Philox counters carry the problem index)
    q = 1234
    prob = ctx.upload(src[:q + 1], dst[:q + 1])
    prob.run(ransac_b200.make_params(thr, c["hypotheses"], sampler=ransac_b200.SAMPLER_PHILOX, seed=4))
    Hs, ms, _ = prob.fetch()
    np.testing.assert_array_equal(Hs[q], H[q])
    np.testing.assert_array_equal(ms[q], mask[q])
    prob.free()


def reference_pnp_inliers(rvec, tvec, K, obj, img, thr):
    """cv::projectPoints (zero distortion) + PnPRansacCallback::computeError restated with NumPy: float64 projection of
    the float32-quantised points in OpenCV's operation order, rounded to float32, float32 squared distance."""
    r = np.asarray(rvec, dtype=np.float64).reshape(3)
    th = np.sqrt(r[0] * r[0] + r[1] * r[1] + r[2] * r[2])
    c, s = np.cos(th), np.sin(th)
    k = r * (1.0 / th)
    rrt = np.outer(k, k)
    rcr = np.array([[0, -k[2], k[1]], [k[2], 0, -k[0]], [-k[1], k[0], 0]])
    R = c * np.eye(3) + (1.0 - c) * rrt + s * rcr
    t = np.asarray(tvec, dtype=np.float64).reshape(3)
    X, Y, Z = (obj[:, i].astype(np.float32).astype(np.float64) for i in range(3))
    x = R[0, 0] * X + R[0, 1] * Y + R[0, 2] * Z + t[0]
    y = R[1, 0] * X + R[1, 1] * Y + R[1, 2] * Z + t[1]
    z = R[2, 0] * X + R[2, 1] * Y + R[2, 2] * Z + t[2]
    iz = np.where(z != 0, 1.0 / np.where(z != 0, z, 1.0), 1.0)
    pu = (x * iz * K[0, 0] + K[0, 2]).astype(np.float32)
    pv = (y * iz * K[1, 1] + K[1, 2]).astype(np.float32)
    dx, dy = img[:, 0].astype(np.float32) - pu, img[:, 1].astype(np.float32) - pv
    return np.nonzero(dx * dx + dy * dy <= np.float32(thr * thr))[0].astype(np.int32)


def test_pnp_large_problem_properties(ctx):
    """The pose model at configs[2] size: 100k points x 100k Philox hypotheses.  The returned inlier list is exactly the
    reference's projectPoints arithmetic applied to the winning minimal model (recomputed with NumPy float64/float32);
    sharded = unsharded; fast arithmetic finds the same consensus up to threshold-borderline points."""
    n, Htot, thr = 100_000, 100_000, 8.0
    P, px, _ = synth.pnp_set(n, 0.5, np.random.default_rng(1898 + 22))
    K = synth.K_1898
    kw = dict(sampler=ransac_b200.SAMPLER_PHILOX, seed=77)
    prob = ctx.upload_pnp(P, px, K)
    prob.run(ransac_b200.make_p_params(thr, Htot, 0.99, arith=ransac_b200.ARITH_EXACT, **kw))
    r0, t0, inl0, i0 = prob.fetch()
    assert i0[0]["status"] == 0 and len(inl0[0]) > 0.4 * n
    ref = reference_pnp_inliers(i0[0]["ransac_rvec"], i0[0]["ransac_tvec"], K, P, px, thr)
    diff = np.setxor1d(ref, inl0[0])
    assert len(diff) <= 2, len(diff)            # NumPy's cos/sin vs the device's: at most a borderline point or two
    keys = [prob.score_shard(ransac_b200.make_p_params(thr, Htot // 4, 0.99, arith=ransac_b200.ARITH_EXACT, hyp_begin=r * (Htot // 4), **kw))
            for r in range(4)]
    best = np.maximum.reduce(keys)
    assert int(best[0]) >> 32 == i0[0]["best_count"] and 0xFFFFFFFF - (int(best[0]) & 0xFFFFFFFF) == i0[0]["best_iter"]
    prob.finish(ransac_b200.make_p_params(thr, Htot // 4, 0.99, arith=ransac_b200.ARITH_EXACT, hyp_begin=0, **kw), best)
    r1, t1, inl1, i1 = prob.fetch()
    np.testing.assert_array_equal(inl1[0], inl0[0])
    np.testing.assert_array_equal(r1, r0)
    np.testing.assert_array_equal(t1, t0)
    prob.run(ransac_b200.make_p_params(thr, Htot, 0.99, arith=ransac_b200.ARITH_FAST, **kw))
    rf, tf, inlf, i_f = prob.fetch()
    assert abs(i_f[0]["best_count"] - i0[0]["best_count"]) <= 5
    if i_f[0]["sample"] == i0[0]["sample"]:
        np.testing.assert_array_equal(inlf[0], inl0[0])      # the returned list always comes from the exact arithmetic
        assert np.abs(rf - r0).max() / np.abs(r0).max() < 1e-5
    prob.free()
