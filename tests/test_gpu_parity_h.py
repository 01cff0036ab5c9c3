"""GPU parity tests, homography path (cv2.findHomography(..., cv2.RANSAC, thr), main_v1.py:312).

Every call goes through the C ABI (ctypes -> libransac_b200.so); the checker is the CPU oracle on the same seeded
inputs.  Integer/index/mask results must be bit-exact; fp64 models from the 4-point solver must be bit-exact; the
refined H is held to 1e-5 relative (north star) and in practice agrees to ~1e-9."""
import numpy as np
import pytest

import ransac_b200
from ransac_b200 import synth

pytestmark = pytest.mark.gpu

REL_H_TOL = 1e-5  # BASELINE.json north_star: "1e-5 relative pose tolerance"


def _quant(a):
    return np.asarray(a, dtype=np.float32)


def _problem(n, outliers, seed, noise=1.0):
    rng = np.random.default_rng(seed)
    s, d, _ = synth.homography_set(n, outliers, rng, noise_px=noise)
    return s, d


def test_rcp_correctly_rounded_exhaustive(ctx):
    bad, tested = ctx.selftest_rcp()
    assert tested > 3_000_000_000  # the fast range covers ~78% of all bit patterns
    assert bad == 0


@pytest.mark.parametrize("n,seed", [(12, 1), (100, 2), (1000, 3), (20000, 4)])
def test_sampler_replays_cv_rng_stream(ctx, oracle, n, seed):
    s, d = _problem(n, 0.4, seed)
    ref = oracle.h_ransac_stage(_quant(s), _quant(d), 3.0, max_iters=300, confidence=1.0)
    got = ctx.sample_cv(_quant(s), _quant(d), 300)
    k = ref["iters"]
    assert k == 300
    np.testing.assert_array_equal(got[:k], ref["samples"])


@pytest.mark.parametrize("n,seed", [(5, 30), (6, 31), (9, 32), (16, 33), (25, 34)])
def test_sampler_under_heavy_rejection(ctx, oracle, n, seed):
    """Lattice points: most 4-subsets contain three collinear points, so checkSubset rejects attempt after attempt and
    duplicate draws are frequent (n = 5: a quarter of the last draws) — the labelling of the RNG stream into attempts and
    iterations (warp-parallel in k_cv_sample_h) must still be OpenCV's, sample for sample."""
    rng = np.random.default_rng(seed)
    side = int(np.ceil(np.sqrt(n)))
    grid = np.array([(100.0 * (i % side), 100.0 * (i // side)) for i in range(n)])
    s = grid + rng.normal(0, 1e-3, grid.shape) * (rng.random((n, 1)) < 0.3)     # a few points slightly off the lattice
    Hgt = np.array([[0.9, 0.05, 10.0], [-0.04, 1.1, 5.0], [1e-5, 2e-5, 1.0]])
    hp = np.c_[s, np.ones(n)] @ Hgt.T
    d = hp[:, :2] / hp[:, 2:]
    bad = rng.permutation(n)[:max(2, n // 3)]                  # outliers keep the adaptive bound from ending the loop
    d[bad] += rng.uniform(20, 60, (len(bad), 2))
    ref = oracle.h_ransac_stage(_quant(s), _quant(d), 3.0, max_iters=200, confidence=1.0)
    got = ctx.sample_cv(_quant(s), _quant(d), 200)
    assert ref["iters"] == 200 and ref["draws"] > 1000         # > 5 RNG outputs per accepted subset
    np.testing.assert_array_equal(got, ref["samples"])


@pytest.mark.parametrize("n", [3, 6, 9])
def test_jacobi_forms_bit_identical_to_oracle(ctx, oracle, n):
    """cv::eigen's Jacobi as calib3d uses it (runKernel's L^T L, cv::solve(DECOMP_EIG) in the LM): the thread, warp and
    packed-shared-memory forms return the oracle's eigenvalues and eigenvectors bit for bit, on well- and ill-conditioned
    matrices (rank-deficient J^T J included)."""
    rng = np.random.default_rng(40 + n)
    mats = []
    for t in range(96):
        rows = n + 3 if t % 3 else n - 1                       # every third matrix is rank deficient
        B = rng.normal(size=(rows, n)) * (10.0 ** rng.uniform(-3, 3, size=(1, n)) if t % 2 else 1.0)
        mats.append(B.T @ B)
    mats.append(np.diag(np.arange(n, 0, -1.0)))                # already diagonal: no rotation at all
    mats.append(np.zeros((n, n)))
    A = np.array(mats)
    ref = [oracle.jacobi(a) for a in A]
    Wr, Vr = np.array([r[0] for r in ref]), np.array([r[1] for r in ref])
    for form in (0, 1, 2) if n == 9 else (0, 1):
        W, V = ctx.jacobi_eig(A, form=form)
        assert np.array_equal(W.view(np.uint64), Wr.view(np.uint64)), f"eigenvalues differ, form {form}"
        assert np.array_equal(V.view(np.uint64), Vr.view(np.uint64)), f"eigenvectors differ, form {form}"


@pytest.mark.parametrize("n,seed", [(12, 5), (500, 6), (5000, 7)])
def test_solver_bit_exact_and_check_subset(ctx, oracle, n, seed):
    s, d = _problem(n, 0.3, seed)
    sq, dq = _quant(s), _quant(d)
    rng = np.random.default_rng(seed)
    idx = np.stack([rng.choice(n, 4, replace=False) for _ in range(400)]).astype(np.int32)
    H, ok, sub = ctx.solve_h4(sq, dq, idx)                                        # thread per solve, shared-memory matrices
    Hw, okw, subw = ctx.solve_h4(sq, dq, idx, solver=2)                           # B2R_SOLVER_EXACT_WARP: one warp per solve
    np.testing.assert_array_equal(H, Hw)
    np.testing.assert_array_equal(ok, okw)
    np.testing.assert_array_equal(sub, subw)
    for k in range(len(idx)):
        Hr = oracle.h_run_kernel(sq[idx[k]], dq[idx[k]])
        assert ok[k] == (Hr is not None)
        if Hr is not None:
            np.testing.assert_array_equal(H[k], Hr)  # fp64 bit-exact
        assert sub[k] == oracle.h_check_subset(sq[idx[k]], dq[idx[k]])


def test_fast_solver_agrees_with_exact(ctx):
    s, d = _problem(2000, 0.3, 12)
    sq, dq = _quant(s), _quant(d)
    rng = np.random.default_rng(12)
    idx = np.stack([rng.choice(2000, 4, replace=False) for _ in range(2000)]).astype(np.int32)
    He, oke, sub = ctx.solve_h4(sq, dq, idx)
    Hf, okf, _ = ctx.solve_h4(sq, dq, idx, solver=ransac_b200.SOLVER_FAST)
    good = oke & okf & sub
    assert good.sum() > 500
    rel = np.abs(He[good] - Hf[good]).max(axis=(1, 2)) / np.abs(He[good]).max(axis=(1, 2))
    # both are fp64 solves of the same 4-point problem; they differ by conditioning only
    assert np.median(rel) < 1e-11 and np.percentile(rel, 99) < 1e-6


@pytest.mark.parametrize("n,n_models,seed", [(12, 100, 8), (1000, 3000, 9), (4097, 5000, 10), (30000, 2048, 11)])
def test_score_exact_counts_bit_exact(ctx, oracle, n, n_models, seed):
    s, d = _problem(n, 0.5, seed)
    sq, dq = _quant(s), _quant(d)
    rng = np.random.default_rng(seed)
    idx = np.stack([rng.choice(n, 4, replace=False) for _ in range(n_models)]).astype(np.int32)
    H, ok, _ = ctx.solve_h4(sq, dq, idx)
    models = H.reshape(-1, 9)[:, :8].astype(np.float32)
    models[~ok] = np.nan
    for thr in (3.0, 75.0):
        thr_sq = np.float32(thr * thr)
        got = ctx.score_h(models, sq, dq, thr_sq, ransac_b200.ARITH_EXACT)
        ref = oracle.h_count_inliers_f32(models, sq, dq, thr_sq)
        np.testing.assert_array_equal(got, ref)
        fast = ctx.score_h(models, sq, dq, thr_sq, ransac_b200.ARITH_FAST)
        # fast mode may move points that sit within rounding of the threshold; never more than a handful
        assert np.abs(fast.astype(np.int64) - ref).max() <= max(3, n // 2000)


def test_score_exact_reciprocal_out_of_range(ctx, oracle):
    """Exact scoring takes the MUFU + Newton reciprocal for a whole batch of points and redoes the batch with the general
    reciprocal when a denominator is zero, infinite, tiny or huge (score_h.cuh): hypotheses that put such denominators
    at the start, the middle and the end of batches, in full tiles and in the tail, must still count bit-exactly."""
    n = 2500
    rng = np.random.default_rng(91)
    src = rng.uniform(0.05, 0.4, (n, 2)).astype(np.float32)
    dst = rng.uniform(0, 2000, (n, 2)).astype(np.float32)
    special = [0, 7, 8, 1023, 1024, 1500, 2047, 2048, 2496, 2499]   # batch and tile boundaries, tail
    src[special] = (1.0, 0.0)          # w = h6 + 1 exactly for these points
    base = np.array([1500, 80, 600, -40, -700, 100, 0.05, -0.02], dtype=np.float32)
    models = np.tile(base, (9, 1))
    models[0, 6] = -1.0                      # w = 0: 1/w = inf
    models[1, 6] = np.float32(3e38)          # w = inf for every point
    models[2, 6] = np.float32(-1.0) + np.float32(2.0 ** -24)  # |w| = 2^-24 .. fine; kept as an in-range control
    models[3, 6:8] = (np.float32(1e35), np.float32(-1e35))    # |w| > 2^101 for most points
    models[4, 6] = -1.0; models[4, 0:3] = 0; models[4, 3:6] = 0   # 0 * inf = NaN at the special points
    models[5] = np.nan
    models[6, 6] = np.float32(-1.0 - 2.0 ** -20)              # tiny negative w at the special points
    models[7, 6:8] = (np.float32(1e-38), np.float32(1e-38))   # denormal products
    for thr_sq in (np.float32(9.0), np.float32(5625.0), np.float32(np.inf)):
        with np.errstate(all="ignore"):
            ref = oracle.h_count_inliers_f32(models, src, dst, thr_sq)
        got = ctx.score_h(models, src, dst, thr_sq, ransac_b200.ARITH_EXACT)
        np.testing.assert_array_equal(got, ref)


def test_score_fast_threshold_edges(ctx, oracle):
    """The fast kernel folds the threshold into its arithmetic (division-free signed margin, score_h.cuh): an infinite
    threshold accepts every point of every finite model and none of a NaN model, a zero threshold accepts (almost)
    nothing, a huge finite one behaves like the exact kernel, and a ragged point count (tile tail) is counted once."""
    n, n_models = 2500, 700   # 2500 = two full 1024-point tiles + a tail of 452
    s, d = _problem(n, 0.5, 77)
    sq, dq = _quant(s), _quant(d)
    rng = np.random.default_rng(77)
    idx = np.stack([rng.choice(n, 4, replace=False) for _ in range(n_models)]).astype(np.int32)
    H, ok, _ = ctx.solve_h4(sq, dq, idx)
    models = H.reshape(-1, 9)[:, :8].astype(np.float32)
    models[~ok] = np.nan
    models[5] = np.nan
    finite = ~np.isnan(models).any(axis=1)
    for thr_sq in (np.float32(np.inf), np.float32(1e30)):
        fast = ctx.score_h(models, sq, dq, thr_sq, ransac_b200.ARITH_FAST)
        ref = oracle.h_count_inliers_f32(models, sq, dq, thr_sq)
        assert (fast[~finite] == 0).all()
        assert np.abs(fast.astype(np.int64) - ref).max() <= 1   # a point on a hypothesis' horizon (w ~ 0) may flip
        assert (fast[finite] >= n - 1).all()
    zero = ctx.score_h(models, sq, dq, np.float32(0.0), ransac_b200.ARITH_FAST)
    ref0 = oracle.h_count_inliers_f32(models, sq, dq, np.float32(0.0))
    assert zero.max() <= 4 and ref0.max() <= 4   # at most the four points the model was fitted to
    for thr in (1e-3, 0.5, 3.0, 1000.0):
        thr_sq = np.float32(thr * thr)
        fast = ctx.score_h(models, sq, dq, thr_sq, ransac_b200.ARITH_FAST)
        ref = oracle.h_count_inliers_f32(models, sq, dq, thr_sq)
        assert np.abs(fast.astype(np.int64) - ref).max() <= 3


@pytest.mark.parametrize("n,outliers,thr,seed", [(5, 0.0, 3.0, 20), (8, 0.2, 3.0, 21), (12, 0.3, 75.0, 22),
                                                 (64, 0.5, 3.0, 23), (300, 0.6, 10.0, 24), (1000, 0.3, 3.0, 25),
                                                 (5000, 0.5, 2.0, 26), (100000, 0.5, 3.0, 27)])
def test_find_homography_matches_oracle(ctx, oracle, n, outliers, thr, seed):
    s, d = _problem(n, outliers, seed)
    H, mask, info = ctx.find_homography(s, d, thr)
    Hr, mr, det = oracle.find_homography(s, d, thr, details=True)
    assert (H is None) == (Hr is None)
    if Hr is None:
        assert mask.sum() == 0
        return
    assert info["iters_run"] == det["iters"]
    assert info["best_count"] == int(det["ransac_mask"].sum())
    np.testing.assert_array_equal(mask, mr)  # final (4.13) mask: identical index set
    assert np.abs(H - Hr).max() / np.abs(Hr).max() < REL_H_TOL
    if n <= 128:   # the reference's problem sizes: refinement summed in OpenCV's order, bit-identical H
        np.testing.assert_array_equal(H, Hr)
        Hp, maskp, _ = ctx.find_homography(s, d, thr, refine=ransac_b200.REFINE_PARALLEL)
        np.testing.assert_array_equal(maskp, mr)
        assert np.abs(Hp - Hr).max() / np.abs(Hr).max() < REL_H_TOL
    # legacy semantics return the RANSAC-stage mask
    _, mask_l, _ = ctx.find_homography(s, d, thr, mask_semantics=ransac_b200.MASK_LEGACY)
    np.testing.assert_array_equal(mask_l.ravel(), det["ransac_mask"])


@pytest.mark.parametrize("max_iters", [1, 2, 63, 64, 65, 191, 192, 193, 447, 448, 449, 1000])
def test_iteration_bound_at_chunk_boundaries(ctx, oracle, max_iters):
    """The replay loop runs in chunks of 64, 128, 256 ... iterations (boundaries 64, 192, 448 ...): a caller's maxIters on,
    just below and just above a boundary, on data whose best model arrives late (60 % outliers: the bound stays above 500
    iterations, the inlier set still improves after iteration 64 and after 192)."""
    s, d = _problem(40, 0.6, 92)
    H, mask, info = ctx.find_homography(s, d, 2.0, max_iters=max_iters, confidence=0.999999)
    Hr, mr, det = oracle.find_homography(s, d, 2.0, max_iters=max_iters, confidence=0.999999, details=True)
    assert (H is None) == (Hr is None)
    assert info["iters_run"] == det["iters"]
    np.testing.assert_array_equal(mask.ravel(), mr.ravel())
    if Hr is not None:
        np.testing.assert_array_equal(H, Hr)          # 40 points: bit-identical refinement


def test_find_homography_four_points_and_errors(ctx, oracle):
    s, d = _problem(4, 0.0, 30)
    H, mask, _ = ctx.find_homography(s, d, 3.0)
    Hr, mr = oracle.find_homography(s, d, 3.0)
    np.testing.assert_array_equal(H, Hr)
    assert mask.ravel().tolist() == [1, 1, 1, 1]
    with pytest.raises(ransac_b200.RansacB200Error):
        ctx.find_homography(s[:3], d[:3], 3.0)
    # all-zero points: no model, zero mask (cv2 returns (None, zeros))
    z = np.zeros((10, 2))
    H, mask, _ = ctx.find_homography(z, z, 3.0)
    assert H is None and mask.sum() == 0


def test_non_finite_input_terminates(ctx, oracle):
    """NaN / inf coordinates: OpenCV's comparisons are all false on NaN, so such subsets pass checkSubset, their models
    count no inliers and the loop runs to its bound.  The GPU path must come back (no hang in the eigen-solver or the
    sampler) with the oracle's mask; with every point NaN there is no model."""
    import time
    s, d = _problem(40, 0.3, 70)
    s2, d2 = s.copy(), d.copy()
    s2[3] = np.nan
    d2[7, 1] = np.inf
    t = time.perf_counter()
    H, mask, info = ctx.find_homography(s2, d2, 3.0, max_iters=500)
    Hr, mr = oracle.find_homography(s2, d2, 3.0, max_iters=500)
    assert (H is None) == (Hr is None)
    np.testing.assert_array_equal(mask.ravel(), mr.ravel())
    assert mask[3] == 0 and mask[7] == 0
    Hn, maskn, _ = ctx.find_homography(np.full((12, 2), np.nan), np.full((12, 2), np.nan), 3.0, max_iters=200)
    assert Hn is None and maskn.sum() == 0
    assert time.perf_counter() - t < 5.0


def test_batch_equals_singles(ctx, oracle):
    rng = np.random.default_rng(40)
    Q, n = 37, 60
    src = np.zeros((Q, n, 2))
    dst = np.zeros((Q, n, 2))
    for q in range(Q):
        src[q], dst[q], _ = synth.homography_set(n, 0.4, rng)
    H, ok, mask, infos = ctx.find_homography_batch(src, dst, 5.0)
    for q in range(Q):
        Hr, mr, det = oracle.find_homography(src[q], dst[q], 5.0, details=True)
        assert ok[q] == (Hr is not None)
        if Hr is None:
            continue
        np.testing.assert_array_equal(mask[q], mr.ravel())
        assert infos[q]["iters_run"] == det["iters"]
        assert np.abs(H[q] - Hr).max() / np.abs(Hr).max() < REL_H_TOL


def test_philox_partition_invariance_and_fast_vs_exact(ctx):
    s, d = _problem(3000, 0.5, 50)
    kw = dict(sampler=ransac_b200.SAMPLER_PHILOX, seed=1234)
    H0, m0, i0 = ctx.find_homography(s, d, 3.0, max_iters=4096, **kw)
    # the same ids scored as two shards give the same winner
    prob = ctx.upload(s, d)
    keys = []
    for begin in (0, 2048):
        p = ransac_b200.make_params(3.0, 2048, hyp_begin=begin, **kw)
        keys.append(prob.score_shard(p))
    best = np.maximum(keys[0], keys[1])
    p = ransac_b200.make_params(3.0, 2048, hyp_begin=0, **kw)
    prob.finish(p, best)
    H1, m1, i1 = prob.fetch()
    np.testing.assert_array_equal(m1[0], m0.ravel())
    np.testing.assert_array_equal(H1[0], H0)
    assert i1[0]["best_count"] == i0["best_count"] and i1[0]["sample"] == i0["sample"]
    # fast arithmetic finds the same winner here and an H within tolerance
    Hf, mf, _ = ctx.find_homography(s, d, 3.0, max_iters=4096, arith=ransac_b200.ARITH_FAST, **kw)
    assert np.abs(Hf - H0).max() / np.abs(H0).max() < REL_H_TOL
    assert (mf != m0).sum() <= 2


@pytest.mark.parametrize("n,seed", [(300, 60), (3000, 61), (4095, 63), (4096, 64), (33000, 65), (40000, 62)])
def test_refine_building_block(ctx, oracle, n, seed):
    """K4's refinement alone (b2r_refine_h): refit on a given inlier set + the 9-parameter LM(10), against the oracle's
    runKernel + LM on the same points.  n >= 4096 takes the cooperative-grid form of the kernel (one point per thread below
    32 768 points, two above), the others one CTA / a cluster; the sums are reduced in a different order than on the CPU,
    hence a tolerance (1e-8) instead of equality."""
    s, d = _problem(n, 0.4, seed)
    sq, dq = _quant(s), _quant(d)
    Hr, mr, det = oracle.find_homography(s, d, 3.0, details=True)
    mask = det["ransac_mask"].astype(np.uint8)
    H, iters = ctx.refine_h(sq, dq, mask, det["ransac_H"])
    inl = mask.astype(bool)
    H0 = oracle.h_run_kernel(sq[inl], dq[inl])
    Href, it_ref = oracle.h_lm_refine(sq[inl], dq[inl], H0)
    if iters != it_ref:
        # The LM stops on a step norm against FLT_EPSILON; where an iterate lands on that edge the count is decided by the
        # last bit of the start (seed 60: the oracle itself runs 5 ... 9 iterations when its own start is perturbed by
        # 1e-16 relative).  The count must lie in the range the oracle shows under such perturbations.
        r = np.random.default_rng(0)
        its = {it_ref} | {oracle.h_lm_refine(sq[inl], dq[inl], H0 * (1 + 1e-15 * r.standard_normal(H0.shape)))[1] for _ in range(24)}
        assert min(its) <= iters <= max(its), (iters, sorted(its))
    assert np.abs(H - Href).max() / np.abs(Href).max() < 1e-8
    assert np.abs(H - Hr).max() / np.abs(Hr).max() < 1e-8          # = what the whole call returns


@pytest.mark.parametrize("n", [5000, 9000, 33000])
def test_batched_cluster_refinement_against_oracle(ctx, oracle, n):
    """A BATCH of problems of >= 4096 points runs one thread-block cluster per problem (2 / 4 / 8 CTAs of 512 threads,
    reductions through distributed shared memory), where a single problem takes the cooperative grid: the refit + LM of
    every problem of the batch against the oracle's runKernel + LM on the RANSAC-stage inlier set the GPU reports."""
    Q = 3
    src, dst = np.zeros((Q, n, 2)), np.zeros((Q, n, 2))
    for q in range(Q):
        src[q], dst[q] = _problem(n, 0.4, 80 + q)
    kw = dict(max_iters=256, sampler=ransac_b200.SAMPLER_PHILOX, seed=11)
    H, ok, mask, infos = ctx.find_homography_batch(src, dst, 3.0, **kw)
    _, _, rmask, _ = ctx.find_homography_batch(src, dst, 3.0, mask_semantics=ransac_b200.MASK_LEGACY, **kw)
    assert ok.all()
    for q in range(Q):
        sq, dq = _quant(src[q]), _quant(dst[q])
        inl = rmask[q].astype(bool)
        assert inl.sum() == infos[q]["best_count"] > n // 4
        Href, _ = oracle.h_lm_refine(sq[inl], dq[inl], oracle.h_run_kernel(sq[inl], dq[inl]))
        assert np.abs(H[q] - Href).max() / np.abs(Href).max() < 1e-8
        assert mask[q].sum() == infos[q]["n_inliers"]          # the count that is now added atomically per CTA
        want = oracle.h_compute_error(Href, sq, dq) <= np.float32(9.0)   # the 4.13 mask of the oracle's refined H
        assert (mask[q].astype(bool) != want).sum() <= 2                  # H agrees to 1e-8: at most a borderline point or two


@pytest.mark.parametrize("solver", ["fast", "exact"])
def test_finalize_is_run_to_run_deterministic(ctx, solver):
    """Every reduction of the finalize kernel has a fixed order (no floating-point atomics), so repeated calls must return
    the same bits: one CTA (300, 1500 points), the cooperative grid (6000, 40 000 points) and clusters of two CTAs (a batch
    of 5000-point problems).  A shared-memory race in the kernel's small sequential part would show up here."""
    arith = ransac_b200.ARITH_FAST if solver == "fast" else ransac_b200.ARITH_EXACT
    sv = ransac_b200.SOLVER_FAST if solver == "fast" else ransac_b200.SOLVER_EXACT
    kw = dict(max_iters=128, sampler=ransac_b200.SAMPLER_PHILOX, seed=5, arith=arith, solver=sv)
    for n in (300, 1500, 6000, 40000):
        s, d = _problem(n, 0.4, 90 + n % 7)
        H0, m0, i0 = ctx.find_homography(s, d, 3.0, **kw)
        for _ in range(8):
            H, m, i = ctx.find_homography(s, d, 3.0, **kw)
            np.testing.assert_array_equal(H, H0)
            np.testing.assert_array_equal(m, m0)
            assert i["n_inliers"] == i0["n_inliers"] == int(m0.sum()) and i["lm_iters"] == i0["lm_iters"]
    src = np.stack([_problem(5000, 0.4, 95 + q)[0] for q in range(3)])
    dst = np.stack([_problem(5000, 0.4, 95 + q)[1] for q in range(3)])
    H0, _, m0, i0 = ctx.find_homography_batch(src, dst, 3.0, **kw)
    for _ in range(8):
        H, _, m, i = ctx.find_homography_batch(src, dst, 3.0, **kw)
        np.testing.assert_array_equal(H, H0)
        np.testing.assert_array_equal(m, m0)
        assert [x["n_inliers"] for x in i] == [int(x.sum()) for x in m0]


@pytest.mark.parametrize("n,seed,hyp_begin", [(12, 70, 0), (500, 71, 0), (500, 72, 2**31 + 12345), (9, 73, 2**32 - 600)])
def test_philox_sampler_matches_restatement(ctx, oracle, n, seed, hyp_begin):
    """North-star kernel 1 (b2r_sample_philox -> k_philox_sample_solve_h): the samples of 600 hypothesis ids equal the
    NumPy restatement (oracle/philox.py: Philox4x32-10 pinned by Random123's known answers, the rejection-free distinct-
    index map, OpenCV's checkSubset from the C oracle, at most 16 attempts), for ids at 0, above 2^31 and up to 2^32 - 1;
    every sample has 4 distinct in-range indices; splitting the id range anywhere returns the same samples; ids beyond
    2^32 are refused (the arg-max key carries the id in 32 bits)."""
    from oracle import philox
    if n == 9:   # lattice: most subsets contain three collinear points, so later attempts are exercised
        s = np.array([(100.0 * (i % 3), 100.0 * (i // 3)) for i in range(9)])
        d = s * 1.1 + 7.0
    else:
        s, d = _problem(n, 0.4, seed)
    sq, dq = _quant(s), _quant(d)
    n_hyp, key = 600, 0x1234_5678_9ABC_DEF0 + seed
    got = ctx.sample_philox(sq, dq, key, hyp_begin, n_hyp)
    want = philox.sample_h(sq, dq, key, hyp_begin, n_hyp, oracle.h_check_subset)
    np.testing.assert_array_equal(got, want)
    ok = got[:, 0] >= 0
    assert ok.sum() >= (n_hyp // 4 if n == 9 else int(0.9 * n_hyp))   # 16 attempts: a sample is missing only when all 16 are rejected
    assert (got[ok] >= 0).all() and (got[ok] < n).all()
    assert all(len(set(row)) == 4 for row in got[ok].tolist())
    if n == 9:
        assert (~ok).sum() == (want[:, 0] < 0).sum()
    for cut in (1, 257, 599):
        a = ctx.sample_philox(sq, dq, key, hyp_begin, cut)
        b = ctx.sample_philox(sq, dq, key, hyp_begin + cut, n_hyp - cut)
        np.testing.assert_array_equal(np.concatenate([a, b]), got)


def test_philox_id_range_is_checked(ctx):
    s, d = _problem(100, 0.3, 80)
    for bad in (dict(hyp_begin=-1, max_iters=10), dict(hyp_begin=2**32 - 5, max_iters=10)):
        with pytest.raises(ransac_b200.RansacB200Error):
            ctx.find_homography(s, d, 3.0, sampler=ransac_b200.SAMPLER_PHILOX, seed=1, **bad)
    H, m, info = ctx.find_homography(s, d, 3.0, sampler=ransac_b200.SAMPLER_PHILOX, seed=1, hyp_begin=2**32 - 10, max_iters=10)
    assert info["iters_run"] == 10
