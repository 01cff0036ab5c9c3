"""Synthetic 2D-3D correspondence sets shaped like the reference's data.

Scene: image 2142x1620 (1898.jpg), K from main_v1.py:870-883, camera at testpro-K.py:234, landmarks uniform in the box
spanned by testpro-K.py:198-211; 1 px Gaussian pixel noise; a given fraction of outliers = uniform pixels.
`rng = np.random.default_rng(1898 + cfg)`."""
import numpy as np

IMAGE_W, IMAGE_H = 2142, 1620
K_1898 = np.array([[240.0 / 127.0 * IMAGE_W, 0.0, 982.666819],
                   [0.0, 240.0 / 178.0 * IMAGE_H, 697.950868],
                   [0.0, 0.0, 1.0]])
CAMERA_ORIGIN = np.array([739424.6, 2888281.18, 770.0])          # testpro-K.py:234
BOX_LO = np.array([738950.0, 2888500.0, 690.0])                  # testpro-K.py:198-211
BOX_HI = np.array([739350.0, 2889050.0, 730.0])

CONFIGS = {
    1: dict(n_points=1_000, outliers=0.30, hypotheses=10_000),
    2: dict(n_points=100_000, outliers=0.50, hypotheses=100_000),
    3: dict(n_points=1_000_000, outliers=0.70, hypotheses=1_000_000),
    4: dict(n_points=2_000, outliers=0.30, hypotheses=2_000, problems=4096),
}


def look_at_pose(origin=CAMERA_ORIGIN, target=None):
    """World->camera rotation R and translation t of a camera at `origin` looking at `target` (z forward, y down)."""
    if target is None:
        target = 0.5 * (BOX_LO + BOX_HI)
    z = target - origin
    z = z / np.linalg.norm(z)
    up = np.array([0.0, 0.0, 1.0])
    x = np.cross(z, up)
    x = x / np.linalg.norm(x)
    y = np.cross(z, x)
    R = np.stack([x, y, z])
    return R, -R @ origin


def pos2_from_camera(pos3d, camera):
    """The reference's per-candidate projection, main_v1.py:306-311: ((z-cz)/(x-cx), (y-cy)/(x-cx)), float64."""
    p = np.asarray(pos3d, dtype=np.float64) - np.asarray(camera, dtype=np.float64)
    return np.stack([p[..., 2] / p[..., 0], p[..., 1] / p[..., 0]], axis=-1)


def pnp_set(n_points, outliers, rng, noise_px=1.0):
    """(pos3d (n,3), pixels (n,2), inlier flags) for path B (cv2.solvePnPRansac, main_v1.py:497)."""
    R, t = look_at_pose()
    P = rng.uniform(BOX_LO, BOX_HI, size=(n_points, 3))
    c = P @ R.T + t
    px = np.stack([K_1898[0, 0] * c[:, 0] / c[:, 2] + K_1898[0, 2], K_1898[1, 1] * c[:, 1] / c[:, 2] + K_1898[1, 2]], axis=1)
    px += rng.normal(0.0, noise_px, size=px.shape)
    out = rng.random(n_points) < outliers
    px[out] = np.stack([rng.uniform(0, IMAGE_W, out.sum()), rng.uniform(0, IMAGE_H, out.sum())], axis=1)
    return P, px, ~out


def _visible_landmarks(n_points, rng, R, t):
    """Landmarks uniform in the box, kept only if they project inside the 2142x1620 image."""
    out = np.zeros((0, 3))
    while len(out) < n_points:
        P = rng.uniform(BOX_LO, BOX_HI, size=(max(1024, 3 * (n_points - len(out))), 3))
        c = P @ R.T + t
        u = K_1898[0, 0] * c[:, 0] / c[:, 2] + K_1898[0, 2]
        v = K_1898[1, 1] * c[:, 1] / c[:, 2] + K_1898[1, 2]
        keep = (c[:, 2] > 1.0) & (u >= 0) & (u < IMAGE_W) & (v >= 0) & (v < IMAGE_H)
        out = np.concatenate([out, P[keep]])
    return out[:n_points]


def homography_set(n_points, outliers, rng, noise_px=1.0):
    """(pos2 (n,2), pixels (n,2), inlier flags) for path A (cv2.findHomography, main_v1.py:312).

    pos2 is the reference's projection of the landmarks from the candidate camera (main_v1.py:306-311), taken at
    the true camera centre, where pos2 -> pixel is exactly a homography; pixels are the pinhole projection plus
    noise; outliers are uniform pixels."""
    R, t = look_at_pose()
    P = _visible_landmarks(n_points, rng, R, t)
    pos2 = pos2_from_camera(P, CAMERA_ORIGIN)
    c = P @ R.T + t
    px = np.stack([K_1898[0, 0] * c[:, 0] / c[:, 2] + K_1898[0, 2], K_1898[1, 1] * c[:, 1] / c[:, 2] + K_1898[1, 2]], axis=1)
    px += rng.normal(0.0, noise_px, size=px.shape)
    out = rng.random(n_points) < outliers
    px[out] = np.stack([rng.uniform(0, IMAGE_W, out.sum()), rng.uniform(0, IMAGE_H, out.sum())], axis=1)
    return pos2, px, ~out


def config_homography(cfg, n_points=None, problems=None):
    """The path-A workload of BASELINE.json configs[cfg] (cfg 1..4)."""
    c = CONFIGS[cfg]
    rng = np.random.default_rng(1898 + cfg)
    n = n_points or c["n_points"]
    if cfg == 4:
        Q = problems or c["problems"]
        src = np.zeros((Q, n, 2))
        dst = np.zeros((Q, n, 2))
        for q in range(Q):
            src[q], dst[q], _ = homography_set(n, c["outliers"], rng)
        return src, dst
    src, dst, _ = homography_set(n, c["outliers"], rng)
    return src, dst
