"""GPU: the reference's OWN scripts — main_v1.find_homographies / find_homography (main_v1.py:254-422), main_v1.
estimate_camera_pose (:468-512), and testpro-K.py from its first to its last line (estimate_camera_orientation, :39-162) —
executed unmodified from the reference's files with `cv2` replaced by ransac_b200.cv2_shim (the injection point of
SURVEY.md §8b), against the golden values recorded from the cv2 4.13.0 binary: err1/err2 of all 458 candidates, the
winning location 180, PnP inliers [0 1 2 3 7 9], the winning intrinsics (#21) and refined poses."""
import numpy as np
import pytest

import reference_harness as rh
from test_reference_scripts_cpu import check_reference_run, fixtures, variant_sweep

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(rh.reference_dir() is None, reason="reference sources not available (build() stages them in baseline/_ref)")]


def test_reference_scripts_on_the_gpu_shim(ctx, tmp_path):
    from ransac_b200 import cv2_shim
    shim = cv2_shim.module()
    calls = {"findHomography": 0, "solvePnPRansac": 0, "solvePnPRefineLM": 0}

    def counted(name, fn):
        def wrapper(*a, **k):
            calls[name] += 1
            return fn(*a, _ctx=ctx, **k)
        return wrapper
    for name in calls:
        setattr(shim, name, counted(name, getattr(shim, name)))
    launches0 = ctx.launch_count()
    m, G = check_reference_run(shim, tmp_path)
    assert m.cv2 is shim
    # 458 candidates + nothing else; one pose in main_v1 + 27 intrinsics in testpro-K; RefineLM once each
    assert calls == {"findHomography": 458, "solvePnPRansac": 1 + 27, "solvePnPRefineLM": 1 + 1}
    assert ctx.launch_count() - launches0 > 458 * 4        # the work ran in this library's kernels


@pytest.mark.parametrize("script", rh.VARIANT_SCRIPTS)
def test_variant_scripts_on_the_gpu_shim(ctx, script, tmp_path):
    """The reference's other pipeline variants (process.py:147-291, testpro.py:293-460, test_pro.py, test02.py): their own
    find_homographies / find_homography over the 458 candidates on the shim — the golden scores of the cv2 binary — and,
    where the variant has one, its own estimate_camera_pose (testpro.py:506-552: solvePnPRansac + solvePnPRefineLM)."""
    from ransac_b200 import cv2_shim
    shim = cv2_shim.module()
    calls = {"findHomography": 0, "solvePnPRansac": 0, "solvePnPRefineLM": 0}

    def counted(name, fn):
        def wrapper(*a, **k):
            calls[name] += 1
            return fn(*a, _ctx=ctx, **k)
        return wrapper
    for name in calls:
        setattr(shim, name, counted(name, getattr(shim, name)))
    m, nm, s, q = variant_sweep(script, shim, tmp_path)
    assert m.cv2 is shim
    run = np.array(s["grids"]) >= m.grid_code_min
    assert calls["findHomography"] == run.sum() >= 400
    np.testing.assert_allclose(nm[:, 0], np.where(run, np.array(s["err1"]), 0), rtol=1e-9)
    np.testing.assert_allclose(nm[:, 1], np.where(run, np.array(s["err2"]), 0), rtol=1e-9)
    if hasattr(m, "estimate_camera_pose"):
        g, k, s, pos3d, pixels, recs = fixtures()
        e = k["estimate_camera_pose"]
        with rh.stubs():
            rvec, tvec, inliers = m.estimate_camera_pose(pos3d, pixels, np.array(e["K"]))
        assert inliers.ravel().tolist() == [0, 1, 2, 3, 7, 9]
        assert np.abs(rvec.ravel() - e["rvec"]).max() / np.abs(e["rvec"]).max() < 1e-5
        assert np.abs(tvec.ravel() - e["tvec"]).max() / np.abs(e["tvec"]).max() < 1e-5
        assert calls["solvePnPRansac"] == 1 and calls["solvePnPRefineLM"] == 1
