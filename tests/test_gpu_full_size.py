"""GPU: BASELINE.json's full-size configurations through size-independent properties (the oracle would need minutes
to hours there): configs[2] = 100k points x 100k hypotheses, configs[3] = 1M points x 1M hypotheses sharded 8 ways,
configs[4] = 4096 problems x 2k points x 2000 hypotheses.

Properties: (a) the returned mask is exactly the reference's fp32 un-fused error formula applied to the
returned H — recomputed here with NumPy float32 arithmetic, which rounds every operation like the reference's scalar
code; (b) the winner of a hypothesis-sharded run equals the winner of the unsharded run (counter-based sampler);
(c) the best count is the maximum of the per-hypothesis counts and the lowest id attaining it; (d) fast arithmetic
finds the same consensus set up to threshold-borderline points."""
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
def test_large_single_problem_properties(ctx, cfg, shards):
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
