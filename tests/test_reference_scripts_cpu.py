"""CPU: the harness that runs the reference's own scripts unmodified (tests/reference_harness.py), with the REAL cv2 —
it must reproduce the golden values recorded from the cv2 binary.  This validates the stand-in modules and the fixtures;
tests/test_gpu_reference_scripts.py then swaps `cv2` for the GPU shim and asserts the same values."""
import json
import os

import numpy as np
import pytest

import reference_harness as rh

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
pytestmark = pytest.mark.skipif(rh.reference_dir() is None, reason="reference sources not available (build() stages them in baseline/_ref)")


def fixtures():
    with open(os.path.join(GOLD, "cv2_golden.json")) as f:
        g = json.load(f)
    with open(os.path.join(GOLD, "cv2_kgrid.json")) as f:
        k = json.load(f)
    s = g["fixture_a_sweep"]
    pos3d, pixels = np.array(s["pos3d"]), np.array(s["pixels"])
    recs = [dict(symbol=str(i), name="", pixel=pixels[i], pos3d=pos3d[i]) for i in range(len(pixels))]
    return g, k, s, pos3d, pixels, recs


def check_reference_run(cv2_module, tmp_path, rel_pose=1e-5):
    """Runs main_v1.find_homographies / estimate_camera_pose and the whole of testpro-K.py from the reference's files and
    checks them against the golden values of the cv2 4.13.0 binary.  cv2_module None: the real OpenCV."""
    g, k, s, pos3d, pixels, recs = fixtures()
    m = rh.import_main_v1(str(tmp_path), cv2_module)
    # the reference's own CSV reader over its own file (pyproj stand-in = Krueger series)
    locs = m.read_camera_locations(os.path.join(rh.reference_dir(), "potential_camera_locations.csv"))
    assert len(locs) == 458
    np.testing.assert_allclose(np.array([c["pos3d"] for c in locs]), np.array(s["loc3ds"]), rtol=0, atol=1e-6)
    assert [c["grid_code"] for c in locs] == s["grids"]
    out = str(tmp_path / "1898.jpg")
    with rh.stubs():
        num_matches = m.find_homographies(recs, locs, None, False, s["thr"], out)          # main_v1.py:254-297, :862
    np.testing.assert_allclose(num_matches[:, 0], np.array(s["err1"]), rtol=1e-9)
    np.testing.assert_allclose(num_matches[:, 1], np.array(s["err2"]), rtol=1e-9)
    err2 = num_matches[:, 1].copy()
    err2[err2 == 0] = 1000000                                                              # main_v1.py:863-866
    assert int(np.argmin(err2)) == s["best_index"] == 180
    with open(out.replace(".jpg", "_location.csv"), encoding="utf-8") as f:
        assert len(f.read().strip().splitlines()) == 459                                   # header + 458 candidates
    e = k["estimate_camera_pose"]
    with rh.stubs():
        rvec, tvec, inliers = m.estimate_camera_pose(pos3d, pixels, np.array(e["K"]))      # main_v1.py:468-512
    assert inliers.ravel().tolist() == [0, 1, 2, 3, 7, 9] == e["inliers"]
    assert np.abs(rvec.ravel() - e["rvec"]).max() / np.abs(e["rvec"]).max() < rel_pose
    assert np.abs(tvec.ravel() - e["tvec"]).max() / np.abs(e["tvec"]).max() < rel_pose
    G = rh.run_testpro_k(str(tmp_path), cv2_module)                                        # testpro-K.py, top to bottom
    assert np.abs(G["R"].ravel() - k["refined_rvec"]).max() / np.abs(k["refined_rvec"]).max() < rel_pose
    assert np.abs(G["T"].ravel() - k["refined_tvec"]).max() / np.abs(k["refined_tvec"]).max() < rel_pose
    return m, G


def test_reference_scripts_with_real_cv2(tmp_path):
    pytest.importorskip("cv2")
    check_reference_run(None, tmp_path, rel_pose=1e-12)


def variant_sweep(script, cv2_module, tmp_path, count=None):
    """find_homographies of one of the reference's other pipeline variants (process.py:147-291, testpro.py:293-460,
    test_pro.py, test02.py), its own function bodies (reference_harness.load_definitions), over the repo's candidates."""
    if not os.path.isfile(os.path.join(rh.reference_dir(), script)):
        pytest.skip(script + " not staged (run __graft_entry__.build() where /root/reference exists)")
    g, k, s, pos3d, pixels, recs = fixtures()
    locs = [dict(grid_code=gc, pos3d=np.array(p)) for gc, p in zip(s["grids"], s["loc3ds"])][:count]
    m = rh.load_definitions(script, str(tmp_path), cv2_module)
    m.output = str(tmp_path / (script + ".png"))            # process.py:179 reads the global its job section sets
    with rh.stubs():
        nm = m.find_homographies(recs, locs, None, False, s["thr"], str(tmp_path / (script + "_out.png")))
    return m, np.asarray(nm), s, len(locs)


@pytest.mark.parametrize("script", rh.VARIANT_SCRIPTS)
def test_variant_scripts_with_real_cv2(script, tmp_path):
    """The stand-ins and the definitions-only loader, validated with the REAL cv2: every variant's sweep reproduces the golden
    scores of the cv2 binary (process.py skips candidates below its own grid_code_min = 7, process.py:398)."""
    pytest.importorskip("cv2")
    m, nm, s, q = variant_sweep(script, None, tmp_path)
    run = np.array(s["grids"][:q]) >= m.grid_code_min
    assert run.sum() >= 400
    np.testing.assert_allclose(nm[:, 0], np.where(run, np.array(s["err1"][:q]), 0), rtol=1e-9)
    np.testing.assert_allclose(nm[:, 1], np.where(run, np.array(s["err2"][:q]), 0), rtol=1e-9)
