"""CPU: the oracle against the golden vectors recorded from the cv2 4.13.0 binary and from the reference's own
debug.log (tests/golden/make_golden.py).  No cv2, no /root/reference, no GPU needed."""
import json
import os

import numpy as np
import pytest

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "cv2_golden.json")
GOLD_LM = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "cv2_lm_blocks.json")
REL_H_TOL = 1e-5   # north star tolerance.  It only applies from 50 RANSAC-stage inliers upwards, where the binary's
#                    J^T r goes through OpenBLAS (cv::gemm hands matrices of >= 100 rows to LAPACK/BLAS, whose summation
#                    order is CPU-kernel specific): observed <= 8e-11.  Below that the refined H is held to EQUALITY.
BLAS_ROWS = 100    # cv::gemm -> cblas_dgemm from this many rows of J (= 2 x inliers), probed on the 4.13.0 binary


@pytest.fixture(scope="module")
def gold():
    with open(GOLD) as f:
        return json.load(f)


def relerr(a, b):
    a, b = np.asarray(a, dtype=np.float64).ravel(), np.asarray(b, dtype=np.float64).ravel()
    return np.abs(a - b).max() / np.abs(b).max()


def test_rng_known_answers(oracle, gold):
    assert oracle.rng_stream(8) == gold["rng_first8"]
    assert [x % 12 for x in oracle.rng_stream(10)] == [9, 4, 8, 9, 3, 8, 9, 8, 5, 6]


def test_jacobi_bit_exact_with_cv2_eigen(oracle, gold):
    for c in gold["eigen9"]:
        W, V = oracle.jacobi(np.array(c["A"]))
        np.testing.assert_array_equal(W, np.array(c["w"]))
        np.testing.assert_array_equal(V, np.array(c["v"]))


def test_four_point_solver_bit_exact(oracle, gold):
    assert len(gold["kernel4"]) >= 50
    for c in gold["kernel4"]:
        H = oracle.h_run_kernel(np.array(c["src"]), np.array(c["dst"]))
        np.testing.assert_array_equal(H, np.array(c["H"]))


def test_find_homography_random_problems(oracle, gold):
    worst, exact = 0.0, 0
    for c in gold["ransac_random"]:
        H, mask, det = oracle.find_homography(np.array(c["src"]), np.array(c["dst"]), c["thr"], details=True)
        assert (H is None) == (c["H"] is None)
        np.testing.assert_array_equal(mask.ravel(), np.array(c["mask"], dtype=np.uint8))
        if H is not None:
            if 2 * int(det["ransac_mask"].sum()) < BLAS_ROWS:
                np.testing.assert_array_equal(H, np.array(c["H"]))      # every bit of the refined H
                exact += 1
            worst = max(worst, relerr(H, c["H"]))
    assert exact >= 20 and worst < 1e-9


def _unhex(a, shape=None):
    v = np.array([float.fromhex(x) for x in a])
    return v.reshape(shape) if shape else v


@pytest.fixture(scope="module")
def gold_lm():
    with open(GOLD_LM) as f:
        return json.load(f)


def test_lm_building_blocks_bit_exact(oracle, gold_lm):
    """cv::LMSolver's linear algebra as the 4.13.0 binary computes it (tests/golden/make_golden_lm.py): the summation
    order of cv::norm (AVX2: fused multiply-adds in a fixed lane order), of cv::gemm's inner products (four interleaved
    partial sums), cv::mulTransposed (one running sum per entry), and the back-substitution of cv::solve /
    cv::invert(DECOMP_EIG) (multiplication by 1/w, threshold 2 eps sum(w))."""
    for c in gold_lm["norm_l2sqr"]:
        assert oracle.cv_norm_l2sqr(_unhex(c["r"])) == float.fromhex(c["out"])
    for c in gold_lm["gemm_atb"]:
        J, r = _unhex(c["J"], (c["rows"], 9)), _unhex(c["r"])
        np.testing.assert_array_equal(oracle.cv_gemm_atb(J, r), _unhex(c["out"]))
    for c in gold_lm["mul_transposed"]:
        J = _unhex(c["J"], (c["rows"], 9))
        A = np.zeros((9, 9))
        for k in range(c["rows"]):
            A += np.outer(J[k], J[k])                                   # one running sum per entry, row order
        np.testing.assert_array_equal(A, _unhex(c["out"], (9, 9)))
    for c in gold_lm["gemm_axpby"]:
        out = oracle.cv_gemm_axpby(_unhex(c["A"], (9, 9)), _unhex(c["d"]), -1.0, _unhex(c["c"]), 2.0)
        np.testing.assert_array_equal(out, _unhex(c["out"]))
    for c in gold_lm["solve_eig"]:
        np.testing.assert_array_equal(oracle.cv_solve_eig(_unhex(c["A"], (9, 9)), _unhex(c["b"])), _unhex(c["out"]))
    for c in gold_lm["invert_eig_diag"]:
        np.testing.assert_array_equal(oracle.cv_invert_eig_diag(_unhex(c["A"], (9, 9))), _unhex(c["out"]))


def test_refit_and_lm_bit_exact_with_cv2(oracle, gold_lm):
    """runKernel on k points + LMSolver(10) == cv2.findHomography(src, dst, 0), every bit of H, on 160 seeded problems
    of 5 ... 39 points (noise 1e-3 ... 10, so early stops, rejected steps and the lambda == 0 branch all occur).  This
    also pins cv::Mat::dot, which the Python binding does not expose (see cv_dot in the oracle)."""
    cases = gold_lm["find_homography_0"]
    assert len(cases) >= 150
    iters = set()
    for c in cases:
        src, dst = _unhex(c["src"], (-1, 2)), _unhex(c["dst"], (-1, 2))
        H0 = oracle.h_run_kernel(src, dst)
        H, it = oracle.h_lm_refine(src, dst, H0)
        iters.add(it)
        np.testing.assert_array_equal(H, _unhex(c["H"], (3, 3)))
    assert len(iters) >= 4          # the stopping rule is exercised at several iteration counts


def test_fixture_a_sweep(oracle, gold):
    """The repo's own data: testpro-K.py:198-225 points x 458 candidate cameras, thr 75 (main_v1.py:862).

    With 12 points and a 75 px threshold most candidates keep only 6-8 inliers, so the answer depends on every detail of
    OpenCV's final Levenberg-Marquardt pass (it is stopped after 10 iterations, far from converged).  With the LM
    restated as the binary runs it — all NINE entries of H are refined, then H is rescaled by 1/H22 — every one of the
    458 candidates agrees with cv2: identical masks, H within 1e-5 (observed max 2.6e-7), err1/err2 and the arg-min."""
    from ransac_b200 import pipeline
    s = gold["fixture_a_sweep"]
    pos3d, pixels, loc3ds = np.array(s["pos3d"]), np.array(s["pixels"]), np.array(s["loc3ds"])
    nm = np.zeros((len(loc3ds), 2))
    worst = 0.0
    for i in range(len(loc3ds)):
        pos2 = pipeline.candidate_pos2(pos3d, loc3ds[i])
        H, mask = oracle.find_homography(pos2, pixels, s["thr"])
        assert H is not None
        np.testing.assert_array_equal(H, np.array(s["H"][i]))          # bit-identical to the cv2 4.13.0 binary
        worst = max(worst, relerr(H, s["H"][i]))
        np.testing.assert_array_equal(mask.ravel(), np.array(s["mask"][i], dtype=np.uint8))
        _, nm[i, 0], nm[i, 1] = pipeline._score(H, mask, pos2, pixels, s["thr"])
    assert worst < REL_H_TOL
    np.testing.assert_allclose(nm[:, 0], np.array(s["err1"]), rtol=1e-5)
    np.testing.assert_allclose(nm[:, 1], np.array(s["err2"]), rtol=1e-5)
    assert pipeline.best_location(nm) == s["best_index"] == 180      # Pointid 181, SURVEY.md Appendix C
    assert abs(nm[180, 1] - 75.212638) < 1e-5


def test_debug_log_known_answers(oracle, gold):
    """The reference's recorded run (thr 120, process.py:374; older OpenCV -> RANSAC-stage "legacy" mask).

    The RANSAC stage (sampler replay, degeneracy tests, 4-point solver, fp32 scoring, termination) reproduces the
    logged mask in all 24 complete blocks, and the logged matrix M = inv(refined H) to the log's printing precision
    and conditioning (<= 1e-3 relative, the same band in which the 4.13 binary reproduces it; SURVEY.md finding 5).
    With OpenCV 4.13 semantics the oracle returns the binary's mask AND the binary's refined H, every bit of it, in
    24/24 blocks (round 1: 4.2e-4 apart on the worst block, until the LM's linear algebra was pinned block by block:
    test_lm_building_blocks_bit_exact)."""
    blocks = gold["debug_log"]
    assert len(blocks) == 24
    rel_M, rel_H = [], []
    for b in blocks:
        pos2, p1 = np.array(b["pos2"]), np.array(b["p1"])
        H, mask_legacy = oracle.find_homography(pos2, p1, 120.0, mask_semantics=1)
        assert mask_legacy.ravel().tolist() == b["logged_mask"]
        M = np.linalg.inv(H)
        M = M * (np.array(b["logged_M"])[2, 2] / M[2, 2])
        rel_M.append(relerr(M, b["logged_M"]))
        H413, mask413 = oracle.find_homography(pos2, p1, 120.0, mask_semantics=0)
        assert mask413.ravel().tolist() == b["cv413_mask"]
        np.testing.assert_array_equal(H413, np.array(b["cv413_H"]))
        rel_H.append(relerr(H413, b["cv413_H"]))
    assert max(rel_M) < 1e-3 and np.median(rel_M) < 1e-6
    assert max(rel_H) == 0.0


def test_project_points_bit_exact(oracle, gold):
    K = np.array(gold["pnp_fixture_a"]["K"])
    for c in gold["project_points"]:
        R = oracle.rodrigues(np.array(c["rvec"]))
        assert np.abs(R - np.array(c["R"])).max() < 1e-15
        proj = oracle.pnp_project_f32(np.array(c["R"]), np.array(c["tvec"]), K, np.array(c["obj"], dtype=np.float32))
        np.testing.assert_array_equal(proj, np.array(c["proj_f32"], dtype=np.float32))


def test_pnp_fixture_a_scoring(oracle, gold):
    """cv2.solvePnPRansac on the repo data returns inliers [0 1 2 3 7 9]; the oracle's PnP
    scoring reproduces cv2.projectPoints for the returned pose, and the replayed 5-point sample stream starts with
    the tuples the survey recorded."""
    p = gold["pnp_fixture_a"]
    assert p["inliers"] == [0, 1, 2, 3, 7, 9]
    s = gold["fixture_a_sweep"]
    obj32 = np.array(s["pos3d"], dtype=np.float32)
    proj = oracle.pnp_project_f32(np.array(p["R"]), np.array(p["tvec"]), np.array(p["K"]), obj32)
    np.testing.assert_array_equal(proj, np.array(p["proj_f32"], dtype=np.float32))
    first = oracle.pnp_sample_stream(12, 6).tolist()
    assert first == [[9, 4, 8, 3, 5], [6, 1, 10, 2, 3], [2, 8, 0, 4, 3], [7, 0, 9, 1, 8], [4, 9, 6, 8, 2], [1, 7, 9, 2, 0]]


def test_epnp_minimal_solver_against_cv2(oracle, gold):
    """PnPRansacCallback::runKernel = solvePnP(5 points, SOLVEPNP_EPNP) + Rodrigues.  With OpenCV's own one-sided
    Jacobi SVD restated (bit-identical to cv2.SVDecomp) the EPnP poses agree with the binary to ~1e-13."""
    K = np.array(gold["pnp_fixture_a"]["K"])
    worst = 0.0
    for c in gold["epnp5"]:
        m = oracle.pnp_minimal_model(np.array(c["obj"], dtype=np.float32), np.array(c["img"], dtype=np.float32), K)
        assert (m is not None) == c["ok"]
        if m is not None:
            worst = max(worst, relerr(m[0], c["rvec"]), relerr(m[1], c["tvec"]))
    assert worst < 1e-9


def test_solve_pnp_ransac_fixture_a(oracle, gold):
    """cv2.solvePnPRansac(..., 5000, 30.0, 0.99) on testpro-K.py:198-225 (main_v1.py:497-502), then solvePnPRefineLM
    (main_v1.py:508): inliers [0 1 2 3 7 9], 145 iterations, both poses within the 1e-5 tolerance (observed 1e-9)."""
    s, p = gold["fixture_a_sweep"], gold["pnp_fixture_a"]
    pos3d, pixels, K = np.array(s["pos3d"]), np.array(s["pixels"]), np.array(p["K"])
    ok, rvec, tvec, inl, det = oracle.solve_pnp_ransac(pos3d, pixels, K, 5000, 30.0, 0.99, details=True)
    assert ok and inl.ravel().tolist() == p["inliers"] == [0, 1, 2, 3, 7, 9]
    assert inl.dtype == np.int32 and inl.shape == (6, 1)
    assert det["iters"] == 145
    assert relerr(rvec, p["rvec"]) < 1e-12 and relerr(tvec, p["tvec"]) < 1e-12
    r2, t2 = oracle.pnp_refine_lm(pos3d[inl.ravel()], pixels[inl.ravel()], K, p["rvec"], p["tvec"])
    assert relerr(r2, p["refined_rvec"]) < 1e-11 and relerr(t2, p["refined_tvec"]) < 1e-11


def test_solve_pnp_ransac_random(oracle, gold):
    K = np.array(gold["pnp_fixture_a"]["K"])
    for c in gold["pnp_ransac_random"]:
        ok, rvec, tvec, inl = oracle.solve_pnp_ransac(np.array(c["obj"]), np.array(c["img"]), K, 5000, c["thr"], 0.99)
        assert ok == c["ok"]
        if not ok:
            continue
        assert inl.ravel().tolist() == c["inliers"]                      # identical inlier index set
        assert relerr(rvec, c["rvec"]) < 1e-12 and relerr(tvec, c["tvec"]) < 1e-12      # north star: 1e-5
        if "refined_rvec" in c:                                            # solvePnPRefineLM on the un-quantised inliers
            idx = inl.ravel()
            r2, t2 = oracle.pnp_refine_lm(np.array(c["obj"])[idx], np.array(c["img"])[idx], K, c["rvec"], c["tvec"])
            assert relerr(r2, c["refined_rvec"]) < 1e-11 and relerr(t2, c["refined_tvec"]) < 1e-11


def test_cv_svd_restatement(oracle):
    """orc_svd is OpenCV's JacobiSVDImpl_: orthogonality and reconstruction here; bit-identity with cv2.SVDecomp was
    checked when the golden file was made (200/200 matrices, including rank-deficient 12x12 Gram matrices)."""
    rng = np.random.default_rng(3)
    for shape in [(3, 3), (6, 4), (12, 12)]:
        A = rng.standard_normal(shape)
        if shape == (12, 12):
            B = rng.standard_normal((10, 12))
            A = B.T @ B
        w, u, vt = oracle.svd(A)
        assert np.all(np.diff(w) <= 1e-12)
        np.testing.assert_allclose((u * w) @ vt, A, atol=1e-9 * max(1.0, np.abs(A).max()))


def test_update_num_iters(oracle):
    assert oracle.update_num_iters(0.995, 0.5, 4, 2000) == 82
    assert oracle.update_num_iters(0.995, 0.0, 4, 2000) == 0
    assert oracle.update_num_iters(0.995, 1.0, 4, 2000) == 2000
    assert oracle.update_num_iters(0.99, 0.5, 5, 5000) == 145   # the PnP run on the repo data executes 145 iterations


def test_intrinsics_grid_testpro_k(oracle):
    """testpro-K.py:58-97 run with the cv2 binary (tests/golden/cv2_kgrid.json): the oracle reproduces, for each of the
    27 camera matrices, success, the inlier index set and the pose; hence the same K wins (f=300, 102x127, 13.41 px)."""
    with open(os.path.join(os.path.dirname(GOLD), "cv2_kgrid.json")) as f:
        kg = json.load(f)
    pos3d, pixels = np.array(kg["pos3d"]), np.array(kg["pixels"])
    used = 0
    for g in kg["grid"]:
        ok, rvec, tvec, inl = oracle.solve_pnp_ransac(pos3d, pixels, np.array(g["K"]), 5000, 30.0, 0.99)
        assert ok == g["ok"]
        if not ok:
            continue
        assert inl.ravel().tolist() == g["inliers"]
        assert relerr(rvec, g["rvec"]) < 1e-12 and relerr(tvec, g["tvec"]) < 1e-12
        used += len(g["inliers"]) >= 6
    assert used == 16 and kg["best"] == 21 and abs(kg["best_error"] - 13.412324744862097) < 1e-9
