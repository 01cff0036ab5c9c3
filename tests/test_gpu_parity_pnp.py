"""GPU parity tests, PnP path: cv2.solvePnPRansac(pos3d, pixels, K, 0, 5000, 30.0, 0.99) (main_v1.py:497-502,
testpro-K.py:72-75) and cv2.solvePnPRefineLM (main_v1.py:508).

Every call goes through the C ABI (ctypes -> libransac_b200.so); the checker is the CPU oracle (pinned against the cv2
binary by tests/test_oracle_golden.py) and the golden files made with cv2 itself.  Sample indices, per-hypothesis inlier
counts, iteration counts and inlier index sets must be identical; poses are held to the north star's 1e-5 relative
tolerance (observed ~1e-9: the refinement is an early-stopped LM whose Jacobian is numerical)."""
import json
import os

import numpy as np
import pytest

import ransac_b200
from ransac_b200 import pipeline, synth

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
REL_POSE_TOL = 1e-5  # BASELINE.json north_star: "1e-5 relative pose tolerance"
K = synth.K_1898


@pytest.fixture(scope="module")
def gold():
    with open(os.path.join(HERE, "golden", "cv2_golden.json")) as f:
        return json.load(f)


@pytest.fixture(scope="module")
def kgrid():
    with open(os.path.join(HERE, "golden", "cv2_kgrid.json")) as f:
        return json.load(f)


def relerr(a, b):
    a, b = np.asarray(a, dtype=np.float64).ravel(), np.asarray(b, dtype=np.float64).ravel()
    return np.abs(a - b).max() / np.abs(b).max()


def _q32(a):
    return np.asarray(a, dtype=np.float32).astype(np.float64)


@pytest.mark.parametrize("n,iters", [(6, 300), (12, 500), (1000, 500), (100000, 200)])
def test_sampler_replays_cv_rng_stream(ctx, oracle, n, iters):
    np.testing.assert_array_equal(ctx.sample_cv_p(n, iters), oracle.pnp_sample_stream(n, iters))


def test_sampler_known_answer(ctx):
    """The first six 5-tuples cv2 draws for the reference's 12 points."""
    assert ctx.sample_cv_p(12, 6).tolist() == [[9, 4, 8, 3, 5], [6, 1, 10, 2, 3], [2, 8, 0, 4, 3], [7, 0, 9, 1, 8],
                                              [4, 9, 6, 8, 2], [1, 7, 9, 2, 0]]


def test_epnp_minimal_models_against_cv2_golden(ctx, gold):
    """PnPRansacCallback::runKernel (EPnP on 5 points + Rodrigues) against cv2.solvePnP(flags=SOLVEPNP_EPNP) itself."""
    Kg = np.array(gold["pnp_fixture_a"]["K"])
    worst = 0.0
    for c in gold["epnp5"]:
        rvec, tvec, R, ok = ctx.pnp_minimal_models(np.array(c["obj"]), np.array(c["img"]), Kg, np.arange(5)[None])
        assert bool(ok[0]) == c["ok"]
        if c["ok"]:
            worst = max(worst, relerr(rvec[0], c["rvec"]), relerr(tvec[0], c["tvec"]))
    assert worst < 1e-9


@pytest.mark.parametrize("n,seed", [(12, 1), (200, 2), (5000, 3)])
def test_epnp_minimal_models_against_oracle(ctx, oracle, n, seed):
    """Same 5-point samples through the device EPnP and the CPU restatement: the translation vectors (no libm call on
    their path) must be bit-identical — the null-space vectors EPnP reads are rounding residue, so anything short of
    the same operation sequence gives different hypotheses; rvec goes through acos and is held to 1e-12."""
    rng = np.random.default_rng(seed)
    P, px, _ = synth.pnp_set(n, 0.3, rng)
    if n == 12:
        P, px = np.array(FIX["pos3d"]), np.array(FIX["pixels"])
    idx = np.stack([rng.choice(len(P), 5, replace=False) for _ in range(300)]).astype(np.int32)
    rvec, tvec, R, ok = ctx.pnp_minimal_models(P, px, K, idx)
    Pq, pq = P.astype(np.float32), px.astype(np.float32)
    same_t = 0
    for k in range(len(idx)):
        m = oracle.pnp_minimal_model(Pq[idx[k]], pq[idx[k]], K)
        assert bool(ok[k]) == (m is not None)
        if m is None:
            continue
        same_t += np.array_equal(tvec[k], m[1])
        assert relerr(tvec[k], m[1]) < 1e-12 and relerr(rvec[k], m[0]) < 1e-12
        assert np.abs(R[k] - oracle.rodrigues(m[0])).max() < 1e-14
    assert same_t >= 0.99 * ok.sum()


@pytest.mark.parametrize("n,n_models,seed", [(12, 64, 4), (1000, 700, 5), (4097, 1500, 6), (30000, 600, 7)])
def test_score_exact_counts_bit_exact(ctx, oracle, n, n_models, seed):
    rng = np.random.default_rng(seed)
    P, px, _ = synth.pnp_set(n, 0.4, rng)
    R0, t0 = synth.look_at_pose()
    models = np.zeros((n_models, 12))
    for m in range(n_models):
        rv = oracle.rodrigues_inv(R0) + rng.normal(0, 2e-3, 3)
        models[m, :9] = oracle.rodrigues(rv).ravel()
        models[m, 9:] = t0 + rng.normal(0, 2.0, 3)
    models[n_models // 2] = np.nan
    for thr in (8.0, 30.0):
        thr_sq = np.float32(thr * thr)
        got = ctx.score_p(models, P, px, K, thr_sq, ransac_b200.ARITH_EXACT)
        ref = np.array([oracle.pnp_count_inliers(m[:9], m[9:], K, P.astype(np.float32), px.astype(np.float32), thr)[0]
                        if np.isfinite(m).all() else 0 for m in models])
        np.testing.assert_array_equal(got, ref)
        fast = ctx.score_p(models, P, px, K, thr_sq, ransac_b200.ARITH_FAST)
        # all-fp32 arithmetic on re-centred points: within 5e-4 px of the exact projection, so only points that sit
        # on the threshold can move
        assert np.abs(fast.astype(np.int64) - ref).max() <= max(3, n // 2000)


def test_score_fast_threshold_edges(ctx, oracle):
    """The fast 3x4 kernel folds the threshold into its arithmetic (division-free signed margin, score_p.cuh): an
    infinite threshold accepts every point of a finite pose and none of a NaN pose; a ragged point count (tile tail)
    and a tiny / huge finite threshold stay within the borderline tolerance of the exact kernel."""
    rng = np.random.default_rng(78)
    n, n_models = 2500, 300
    P, px, _ = synth.pnp_set(n, 0.4, rng)
    R0, t0 = synth.look_at_pose()
    models = np.zeros((n_models, 12))
    for m in range(n_models):
        rv = oracle.rodrigues_inv(R0) + rng.normal(0, 2e-3, 3)
        models[m, :9] = oracle.rodrigues(rv).ravel()
        models[m, 9:] = t0 + rng.normal(0, 2.0, 3)
    models[7] = np.nan
    finite = np.isfinite(models).all(axis=1)
    fast = ctx.score_p(models, P, px, K, np.float32(np.inf), ransac_b200.ARITH_FAST)
    assert (fast[~finite] == 0).all() and (fast[finite] == n).all()
    for thr in (0.05, 8.0, 3000.0):
        thr_sq = np.float32(thr * thr)
        exact = ctx.score_p(models, P, px, K, thr_sq, ransac_b200.ARITH_EXACT)
        fast = ctx.score_p(models, P, px, K, thr_sq, ransac_b200.ARITH_FAST)
        assert np.abs(fast.astype(np.int64) - exact).max() <= 3
    zero = ctx.score_p(models, P, px, K, np.float32(0.0), ransac_b200.ARITH_FAST)
    assert zero.max() <= 1


FIX = {}


@pytest.fixture(autouse=True, scope="module")
def _load_fixture(gold):
    FIX.update(pos3d=gold["fixture_a_sweep"]["pos3d"], pixels=gold["fixture_a_sweep"]["pixels"])


def test_solve_pnp_ransac_reference_data(ctx, oracle, gold):
    """The repo's own 12 correspondences (testpro-K.py:198-225), K of main_v1.py:870-883: inliers [0 1 2 3 7 9], 145
    iterations, winner = iteration 5 with sample (1,7,9,2,0); pose against the cv2 binary."""
    p = gold["pnp_fixture_a"]
    pos3d, pixels, Kg = np.array(FIX["pos3d"]), np.array(FIX["pixels"]), np.array(p["K"])
    ok, rvec, tvec, inl, info = ctx.solve_pnp_ransac(pos3d, pixels, Kg, 5000, 30.0, 0.99)
    assert ok and inl.dtype == np.int32 and inl.shape == (6, 1)
    assert inl.ravel().tolist() == p["inliers"] == [0, 1, 2, 3, 7, 9]
    assert info["iters_run"] == 145 and info["best_iter"] == 5 and info["sample"] == [1, 7, 9, 2, 0] and info["best_count"] == 6
    assert relerr(rvec, p["rvec"]) < REL_POSE_TOL and relerr(tvec, p["tvec"]) < REL_POSE_TOL
    _, r_o, t_o, _, det = oracle.solve_pnp_ransac(pos3d, pixels, Kg, 5000, 30.0, 0.99, details=True)
    assert relerr(info["ransac_rvec"], det["ransac_rvec"]) < 1e-12 and relerr(info["ransac_tvec"], det["ransac_tvec"]) < 1e-12
    assert relerr(rvec, r_o) < 1e-7 and relerr(tvec, t_o) < 1e-7
    r2, t2, _ = ctx.solve_pnp_refine_lm(pos3d[inl.ravel()], pixels[inl.ravel()], Kg, p["rvec"], p["tvec"])
    assert relerr(r2, p["refined_rvec"]) < REL_POSE_TOL and relerr(t2, p["refined_tvec"]) < REL_POSE_TOL


def test_solve_pnp_ransac_golden_random(ctx, gold):
    """24 seeded synthetic problems (5 ... 2000 points) whose answers were produced by cv2.solvePnPRansac itself."""
    Kg = np.array(gold["pnp_fixture_a"]["K"])
    worst = 0.0
    for c in gold["pnp_ransac_random"]:
        obj, img = np.array(c["obj"]), np.array(c["img"])
        ok, rvec, tvec, inl, _ = ctx.solve_pnp_ransac(obj, img, Kg, 5000, c["thr"], 0.99)
        assert ok == c["ok"]
        if not ok:
            assert inl is None
            continue
        assert inl.ravel().tolist() == c["inliers"]                       # identical inlier index set
        worst = max(worst, relerr(rvec, c["rvec"]), relerr(tvec, c["tvec"]))
        if "refined_rvec" in c:
            r2, t2, _ = ctx.solve_pnp_refine_lm(obj[inl.ravel()], img[inl.ravel()], Kg, c["rvec"], c["tvec"])
            worst = max(worst, relerr(r2, c["refined_rvec"]), relerr(t2, c["refined_tvec"]))
    assert worst < REL_POSE_TOL


@pytest.mark.parametrize("n,outliers,thr,seed", [(6, 0.0, 8.0, 30), (40, 0.3, 8.0, 31), (1000, 0.5, 8.0, 32),
                                                 (20000, 0.5, 4.0, 33), (100000, 0.3, 8.0, 34)])
def test_solve_pnp_ransac_matches_oracle(ctx, oracle, n, outliers, thr, seed):
    rng = np.random.default_rng(seed)
    P, px, _ = synth.pnp_set(n, outliers, rng)
    ok, rvec, tvec, inl, info = ctx.solve_pnp_ransac(P, px, K, 5000, thr, 0.99)
    ok_o, r_o, t_o, inl_o, det = oracle.solve_pnp_ransac(P, px, K, 5000, thr, 0.99, details=True)
    assert ok == ok_o
    if not ok:
        return
    assert info["iters_run"] == det["iters"]
    np.testing.assert_array_equal(inl, inl_o)
    assert relerr(rvec, r_o) < REL_POSE_TOL and relerr(tvec, t_o) < REL_POSE_TOL


@pytest.mark.parametrize("iterations", [1, 255, 256, 257, 767, 768, 769, 1500])
def test_iteration_count_at_chunk_boundaries(ctx, oracle, iterations):
    """The PnP replay loop runs in chunks of 256, 512, 1024 ... iterations (boundaries 256, 768, 1792 ...): iterationsCount on,
    just below and just above a boundary, with 70 % outliers so that the bound stays high."""
    rng = np.random.default_rng(77)
    P, px, _ = synth.pnp_set(60, 0.7, rng)
    ok, rvec, tvec, inl, info = ctx.solve_pnp_ransac(P, px, K, iterations, 4.0, 0.999999)
    ok_o, r_o, t_o, inl_o, det = oracle.solve_pnp_ransac(P, px, K, iterations, 4.0, 0.999999, details=True)
    assert ok == ok_o
    if ok:
        assert info["iters_run"] == det["iters"]
        np.testing.assert_array_equal(inl, inl_o)
        assert relerr(rvec, r_o) < REL_POSE_TOL and relerr(tvec, t_o) < REL_POSE_TOL


def test_per_iteration_counts_match_oracle(ctx, oracle):
    """Every executed RANSAC iteration scores the same number of inliers as the CPU path (replayed samples, EPnP
    models, exact scoring): checked through the building blocks on the reference's data."""
    pos3d, pixels = np.array(FIX["pos3d"]), np.array(FIX["pixels"])
    st = oracle.pnp_ransac_stage(pos3d.astype(np.float32), pixels.astype(np.float32), K, 5000, 30.0, 0.99)
    idx = ctx.sample_cv_p(12, st["iters"])
    rvec, tvec, R, ok = ctx.pnp_minimal_models(pos3d, pixels, K, idx)
    models = np.concatenate([R.reshape(-1, 9), tvec], axis=1)
    models[~ok] = np.nan
    counts = ctx.score_p(models, pos3d, pixels, K, np.float32(900.0), ransac_b200.ARITH_EXACT)
    ref = np.where(st["counts"] < 0, 0, st["counts"])
    np.testing.assert_array_equal(counts, ref)


def test_intrinsics_grid_batch(ctx, kgrid):
    """estimate_camera_orientation (testpro-K.py:39-162): the 27 camera matrices as ONE batched call over the shared
    points, against what the cv2 binary returned per K; same winning K, same refined pose."""
    pos3d, pixels = np.array(kgrid["pos3d"]), np.array(kgrid["pixels"])
    rvec, tvec, det = pipeline.estimate_camera_orientation(pos3d, pixels, kgrid["focal_lengths"],
                                                           [tuple(s) for s in kgrid["sensor_sizes"]], tuple(kgrid["image_size"]),
                                                           np.array(kgrid["known_camera_origin"]), ctx=ctx, return_details=True)
    Ks, _ = pipeline.intrinsics_grid(kgrid["focal_lengths"], [tuple(s) for s in kgrid["sensor_sizes"]], tuple(kgrid["image_size"]))
    ok, rv, tv, inl, infos = ctx.solve_pnp_ransac_batch(pos3d, pixels, Ks, 5000, 30.0, 0.99)
    for q, g in enumerate(kgrid["grid"]):
        np.testing.assert_allclose(Ks[q], np.array(g["K"]), rtol=0, atol=0)
        assert bool(ok[q]) == g["ok"]
        if not g["ok"]:
            continue
        assert inl[q].tolist() == g["inliers"]
        assert relerr(rv[q], g["rvec"]) < REL_POSE_TOL and relerr(tv[q], g["tvec"]) < REL_POSE_TOL
        if g["used"]:
            assert abs(infos[q]["mean_inlier_err"] - g["mean_error"]) < 1e-4   # north star: 1e-4 px reprojection
    assert det["best"] == kgrid["best"] == 21
    assert abs(det["best_error"] - kgrid["best_error"]) < 1e-4
    assert relerr(rvec, kgrid["refined_rvec"]) < REL_POSE_TOL and relerr(tvec, kgrid["refined_tvec"]) < REL_POSE_TOL


def test_estimate_camera_pose_chain(ctx, kgrid):
    """main_v1.py:468-512: solvePnPRansac -> `< 6 inliers` gate -> solvePnPRefineLM, against the cv2 binary."""
    e = kgrid["estimate_camera_pose"]
    rvec, tvec, inl = pipeline.estimate_camera_pose(np.array(kgrid["pos3d"]), np.array(kgrid["pixels"]), np.array(e["K"]), ctx=ctx)
    assert inl.ravel().tolist() == e["inliers"]
    assert relerr(rvec, e["rvec"]) < REL_POSE_TOL and relerr(tvec, e["tvec"]) < REL_POSE_TOL


def test_cv2_shim_pnp_signature(ctx, gold):
    from ransac_b200 import cv2_shim
    shim = cv2_shim.module()
    p = gold["pnp_fixture_a"]
    pos3d, pixels, Kg = np.array(FIX["pos3d"]), np.array(FIX["pixels"]), np.array(p["K"])
    ok, rvec, tvec, inl = shim.solvePnPRansac(pos3d, pixels, Kg, np.zeros((4, 1)), iterationsCount=5000, reprojectionError=30.0,
                                              confidence=0.99)
    assert ok is True and rvec.shape == (3, 1) and tvec.shape == (3, 1) and rvec.dtype == np.float64
    assert inl.dtype == np.int32 and inl.shape == (6, 1) and inl.ravel().tolist() == p["inliers"]
    r2, t2 = shim.solvePnPRefineLM(pos3d[inl], pixels[inl], Kg, np.zeros((4, 1)), rvec, tvec)   # (k,1,3) like main_v1.py:508
    assert r2.shape == (3, 1) and relerr(r2, p["refined_rvec"]) < REL_POSE_TOL and relerr(t2, p["refined_tvec"]) < REL_POSE_TOL


def test_edge_cases(ctx, oracle):
    rng = np.random.default_rng(60)
    P, px, _ = synth.pnp_set(5, 0.0, rng, noise_px=0.0)                   # n == modelPoints: one solve, all must be inliers
    ok, rvec, tvec, inl, info = ctx.solve_pnp_ransac(P, px, K, 5000, 8.0, 0.99)
    ok_o, r_o, t_o, inl_o = oracle.solve_pnp_ransac(P, px, K, 5000, 8.0, 0.99)
    assert ok == ok_o
    if ok:
        np.testing.assert_array_equal(inl, inl_o)
        assert relerr(rvec, r_o) < REL_POSE_TOL
    with pytest.raises(ransac_b200.RansacB200Error):
        ctx.solve_pnp_ransac(P[:3], px[:3], K)                          # cv2 raises below 4 points
    with pytest.raises(ransac_b200.RansacB200Error):
        ctx.solve_pnp_ransac(P[:4], px[:4], K)                          # P3P branch: not on the reference's path
    Pz, pz, _ = synth.pnp_set(30, 1.0, rng)                             # pure outliers: retval False
    ok, _, _, inl, _ = ctx.solve_pnp_ransac(Pz, pz, K, 200, 1.0, 0.99)
    ok_o = oracle.solve_pnp_ransac(Pz, pz, K, 200, 1.0, 0.99)[0]
    assert ok == ok_o and (ok or inl is None)


def test_philox_partition_invariance_and_fast_vs_exact(ctx):
    rng = np.random.default_rng(70)
    P, px, _ = synth.pnp_set(3000, 0.5, rng)
    kw = dict(sampler=ransac_b200.SAMPLER_PHILOX, seed=99)
    ok0, r0, t0, inl0, i0 = ctx.solve_pnp_ransac(P, px, K, 4096, 8.0, 0.99, **kw)
    assert ok0 and len(inl0) > 1000
    prob = ctx.upload_pnp(P, px, K)
    keys = []
    for begin in (0, 2048):
        keys.append(prob.score_shard(ransac_b200.make_p_params(8.0, 2048, 0.99, hyp_begin=begin, **kw)))
    best = np.maximum(keys[0], keys[1])
    prob.finish(ransac_b200.make_p_params(8.0, 2048, 0.99, hyp_begin=0, **kw), best)
    r1, t1, inl1, i1 = prob.fetch()
    np.testing.assert_array_equal(inl1[0], inl0.ravel())
    np.testing.assert_array_equal(r1[0], r0.ravel())
    assert i1[0]["sample"] == i0["sample"] and i1[0]["best_count"] == i0["best_count"]
    okf, rf, tf, inlf, i_f = ctx.solve_pnp_ransac(P, px, K, 4096, 8.0, 0.99, arith=ransac_b200.ARITH_FAST, **kw)
    assert okf and abs(i_f["best_count"] - i0["best_count"]) <= 3
    if i_f["sample"] == i0["sample"]:
        assert relerr(rf, r0) < REL_POSE_TOL and relerr(tf, t0) < REL_POSE_TOL


def test_fast_minimal_solver(ctx):
    """B2R_SOLVER_FAST (depth-parametrised 5-point solver + Gauss-Newton, the throughput path): exact on noise-free
    samples; in a Philox RANSAC on noisy data it finds the same consensus as the EPnP-based run up to a few points."""
    R0, t0 = synth.look_at_pose()
    rng = np.random.default_rng(80)
    P, px, _ = synth.pnp_set(400, 0.0, rng, noise_px=0.0)
    # local coordinates: OpenCV's float32 input quantisation moves raw UTM coordinates by up to 0.25 m (there is then no
    # exact pose); relative to the corner of the landmark box it moves them by 3e-5 m
    P = P - synth.BOX_LO
    t0 = t0 + R0 @ synth.BOX_LO
    idx = np.stack([rng.choice(400, 5, replace=False) for _ in range(500)]).astype(np.int32)
    rvec, tvec, R, ok = ctx.pnp_minimal_models(P, px, K, idx, solver=ransac_b200.SOLVER_FAST)
    assert ok.mean() > 0.99
    errR = np.abs(R[ok] - R0).max(axis=(1, 2))
    errt = np.abs(tvec[ok] - t0).max(axis=1) / np.abs(t0).max()
    assert np.median(errR) < 1e-6 and np.median(errt) < 1e-6 and np.percentile(errR, 95) < 1e-4
    P, px, _ = synth.pnp_set(5000, 0.5, rng)
    kw = dict(sampler=ransac_b200.SAMPLER_PHILOX, seed=5, arith=ransac_b200.ARITH_FAST)
    ok_e, r_e, t_e, inl_e, i_e = ctx.solve_pnp_ransac(P, px, K, 8192, 8.0, 0.99, solver=ransac_b200.SOLVER_EXACT, **kw)
    ok_f, r_f, t_f, inl_f, i_f = ctx.solve_pnp_ransac(P, px, K, 8192, 8.0, 0.99, solver=ransac_b200.SOLVER_FAST, **kw)
    assert ok_e and ok_f
    assert abs(len(inl_f) - len(inl_e)) <= 0.02 * len(inl_e)
    # both poses are early-stopped LM refinements (OpenCV's criterion fires after ~2 iterations on UTM-scale translations)
    # of different minimal models on nearly the same inliers: compare the camera centres (the scene is ~600 m away)
    o_e = -pipeline.rodrigues(r_e).T @ t_e.ravel()
    o_f = -pipeline.rodrigues(r_f).T @ t_f.ravel()
    assert np.linalg.norm(o_e - synth.CAMERA_ORIGIN) < 10.0 and np.linalg.norm(o_f - synth.CAMERA_ORIGIN) < 10.0
