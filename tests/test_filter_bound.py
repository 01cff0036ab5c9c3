"""The error bounds of the filtered exact predicates (csrc/score_h_filt.cuh, csrc/score_p_filt.cuh), checked on the CPU in exact
arithmetic.

The kernel accepts the sign of its division-free fp32 FMA margin when |margin| >= 2.0 after scaling the hypothesis by
(2.002 / B)^1/2.  Here that margin is re-computed with correctly rounded fp32 FMAs built from Python fractions (no GPU, no
libm), the per-hypothesis scale from the same fp32 operations as k3_filter_const(), and OpenCV's comparison with numpy's
individually rounded float32 operations (HomographyEstimatorCallback::computeError; reference call site
main_v1.py:312).  Pixels are placed so that the squared error lands within 1e-8 .. 1e-2 (relative) of the threshold, on
hypotheses with the cancellation of the reference's pos2 coordinates, with vanishing denominators, and with huge and tiny
coefficients.  Claim under test: decided  =>  (margin < 0) == (err_cv <= thr).  The slack of the bound is reported.
The second test does the same for the 3x4 kernel (cv::projectPoints' fp64 projection + the fp32 error; reference call
site main_v1.py:497-502), its fp64 set-up, its fp32 margin on re-centred points and its depth guard."""
import math
from fractions import Fraction

import numpy as np

F32 = np.float32
U = F32(2.0 ** -24)


def rn32(x):
    """Correctly rounded (nearest-even) binary32 of a Fraction, normal range."""
    if x == 0:
        return F32(0.0)
    sign = -1 if x < 0 else 1
    x = abs(x)
    e = x.numerator.bit_length() - x.denominator.bit_length()
    if Fraction(2) ** e > x:
        e -= 1                      # 2^e <= x < 2^(e+1)
    scaled = x / Fraction(2) ** (e - 23)            # in [2^23, 2^24)
    n, r = divmod(scaled.numerator, scaled.denominator)
    twice = 2 * r
    if twice > scaled.denominator or (twice == scaled.denominator and (n & 1)):
        n += 1
    assert -126 <= e <= 126
    return F32(sign * math.ldexp(n, e - 23))


def fr(x):
    return Fraction(float(x))


def fma32(a, b, c):
    return rn32(fr(a) * fr(b) + fr(c))


def filter_const(h, Xm, Ym, Um, Vm, s):
    """k3_filter_const() of csrc/score_h_filt.cuh, operation for operation (float32, nothing fused)."""
    up, big = F32(1.0) + F32(2.0 ** -18), F32(2.0 ** 40)
    a = np.abs(h)
    Ax = (a[0] * Xm + a[1] * Ym + a[2]) * up
    Ay = (a[3] * Xm + a[4] * Ym + a[5]) * up
    Aw = (a[6] * Xm + a[7] * Ym + F32(1.0)) * up
    D = ((F32(9.5) * U * s) * (Ax + Ay) * up + ((F32(8.5) * U * s) * (Um + Vm) * up + F32(16.0) * U) * Aw) * up + F32(2.0 ** -60)
    B = (F32(2.0) * (Aw * up) * D + D * D) * (F32(1.0) + F32(2.0 ** -9))
    # rsqrtf is within 2 ulp of 1/sqrt: take the UNSAFE end (a larger scale decides more)
    kappa = F32(F32(1.0 / math.sqrt(float(B))) * (F32(1.0) + F32(3 * 2.0 ** -23)) * F32(1.4150))
    return kappa, bool(Ax > big or Ay > big or Aw > big), float(B)


def margin(h, kappa, s, X, Y, nu_s, nv_s):
    ksa = F32(kappa * s)
    hs = [F32(h[k] * ksa) for k in range(6)] + [F32(h[6] * kappa), F32(h[7] * kappa), kappa]
    w = fma32(hs[6], X, fma32(hs[7], Y, hs[8]))
    sx = fma32(hs[0], X, fma32(hs[1], Y, hs[2]))
    sy = fma32(hs[3], X, fma32(hs[4], Y, hs[5]))
    a = fma32(w, nu_s, sx)
    b = fma32(w, nv_s, sy)
    t = rn32(fr(w) * fr(w))
    return fma32(a, a, fma32(b, b, -t))


def cv_error(h, X, Y, u, v):
    one = F32(1.0)
    with np.errstate(all="ignore"):
        ww = one / ((h[6] * X + h[7] * Y) + one)
        dx = ((h[0] * X + h[1] * Y) + h[2]) * ww - u
        dy = ((h[3] * X + h[4] * Y) + h[5]) * ww - v
        return dx * dx + dy * dy


def _hypotheses(rng):
    # the bench's geometry: pos2 coordinates in (0.08..1, -10..-0.5), coefficients in the thousands, heavy cancellation
    base = np.array([3600.0, 40.0, 900.0, 250.0, -330.0, -200.0, 0.9, 0.02])
    out = []
    for k in range(60):
        out.append(base * (1 + rng.normal(0, [1e-5, 1e-3, 0.05][k % 3], 8)))
    for k in range(20):      # denominators that vanish inside the point range
        hh = base * (1 + rng.normal(0, 0.3, 8))
        hh[6:8] = (-1.0 / rng.uniform(0.2, 0.9), rng.normal(0, 0.02))
        out.append(hh)
    for k in range(10):
        out.append(base * 10.0 ** rng.uniform(-6, 6) * (1 + rng.normal(0, 0.1, 8)))
    for k in range(10):
        hh = rng.normal(0, 1, 8) * 10.0 ** rng.uniform(-3, 5, 8)
        out.append(hh)
    return np.array(out).astype(F32)


def test_decided_margins_agree_with_opencv():
    rng = np.random.default_rng(2026)
    hyps = _hypotheses(rng)
    n_pts = 48
    checked = decided = wrong_sign = 0
    worst = 0.0     # largest |scaled margin| / 2.0 among evaluations whose margin sign differs from OpenCV's decision
    for thr in (F32(9.0), F32(1.0), F32(5625.0)):
        for s_bias in (0, 1, -1):
            s = F32(F32(1.0 / math.sqrt(float(thr))) * (F32(1.0) + F32(s_bias * 2.0 ** -22)))   # rsqrtf: within 2 ulp
            for h in hyps:
                X = rng.uniform(0.08, 1.0, n_pts).astype(F32)
                Y = rng.uniform(-10.0, -0.5, n_pts).astype(F32)
                hd = h.astype(np.float64)
                w = hd[6] * X + hd[7] * Y + 1.0
                pu = (hd[0] * X + hd[1] * Y + hd[2]) / w
                pv = (hd[3] * X + hd[4] * Y + hd[5]) / w
                ang = rng.uniform(0, 2 * np.pi, n_pts)
                rel = 10.0 ** rng.uniform(-8, -2, n_pts) * rng.choice([-1.0, 1.0], n_pts)
                rad = math.sqrt(float(thr)) * (1.0 + rel)
                u = (pu + rad * np.cos(ang)).astype(F32)
                v = (pv + rad * np.sin(ang)).astype(F32)
                ok = np.isfinite(u) & np.isfinite(v) & (np.abs(u) < 1e6) & (np.abs(v) < 1e6)
                if ok.sum() < 8:
                    continue
                X, Y, u, v = X[ok], Y[ok], u[ok], v[ok]
                kappa, forced, _ = filter_const(h, np.abs(X).max(), np.abs(Y).max(), np.abs(u).max(), np.abs(v).max(), s)
                if forced or not np.isfinite(kappa):
                    continue
                e = cv_error(h, X, Y, u, v)
                for i in range(len(X)):
                    m = margin(h, kappa, s, X[i], Y[i], F32(-u[i] * s), F32(-v[i] * s))
                    inl_cv = bool(e[i] <= thr)
                    checked += 1
                    if (m < 0) != inl_cv:
                        wrong_sign += 1
                        worst = max(worst, abs(float(m)) / 2.0)
                    if abs(float(m)) >= 2.0:
                        decided += 1
                        assert (m < 0) == inl_cv, (h, X[i], Y[i], u[i], v[i], float(m), float(e[i]), float(thr))
    # the sample must exercise both sides: decided and undecided evaluations, and margins whose sign is NOT OpenCV's answer
    assert checked > 20000 and decided > checked // 10 and decided < checked
    assert wrong_sign > 20
    assert worst < 1.0
    print(f"checked {checked}, decided {decided}, margin sign != cv decision {wrong_sign}, worst |margin|/band {worst:.3f}")


# ---- 3x4 models: csrc/score_p_filt.cuh ---------------------------------------------------------------------------------------

def p_filter_const(m, K4, c, Xm, Um, Vm, s):
    """The per-hypothesis set-up of k3_score_p_filt, operation for operation: P = K [R | R c + t] in fp64, the magnitudes, D, B,
    kappa, zeta and the depth guard's factor in fp32.  m: 12 doubles (R | t); Xm: the tile's max |X_j - c_j| (3 floats)."""
    fx, fy, cx, cy = (float(v) for v in K4)
    up, uu, big = F32(1.0) + F32(2.0 ** -18), U, F32(2.0 ** 40)
    tc = [((m[3 * i] * c[0] + m[3 * i + 1] * c[1]) + m[3 * i + 2] * c[2]) + m[9 + i] for i in range(3)]
    P = np.zeros(12, dtype=F32)
    for k in range(3):
        P[k] = F32(fx * m[k] + cx * m[6 + k])
        P[4 + k] = F32(fy * m[3 + k] + cy * m[6 + k])
        P[8 + k] = F32(m[6 + k])
    P[3], P[7], P[11] = F32(fx * tc[0] + cx * tc[2]), F32(fy * tc[1] + cy * tc[2]), F32(tc[2])
    aP = np.abs(P)
    Ax = (aP[0] * Xm[0] + aP[1] * Xm[1] + aP[2] * Xm[2] + aP[3]) * up
    Ay = (aP[4] * Xm[0] + aP[5] * Xm[1] + aP[6] * Xm[2] + aP[7]) * up
    Aw = (aP[8] * Xm[0] + aP[9] * Xm[1] + aP[10] * Xm[2] + aP[11]) * up
    a = [F32(abs(float(c[j])) * (1 + 1e-6)) + Xm[j] for j in range(3)]
    am = np.abs(np.asarray(m, dtype=np.float64)).astype(F32)
    Gx = (am[0] * a[0] + am[1] * a[1] + am[2] * a[2] + am[9]) * up
    Gy = (am[3] * a[0] + am[4] * a[1] + am[5] * a[2] + am[10]) * up
    Gz = (am[6] * a[0] + am[7] * a[1] + am[8] * a[2] + am[11]) * up
    afx, afy = F32(abs(fx)) * up, F32(abs(fy)) * up
    acx, acy = F32(abs(cx)) * up + Um, F32(abs(cy)) * up + Vm
    e50 = F32(2.0 ** -50)
    E64 = e50 * ((afx * Gx + acx * Gz) + (afy * Gy + acy * Gz)) * up
    yw = e50 * Gz * up
    D = (s * (F32(7.2) * uu * (Ax + Ay) + F32(10.5) * uu * (Um + Vm) * Aw + F32(4.2) * E64) * up + F32(19.2) * uu * Aw + yw) * up + F32(2.0 ** -60)
    B = (F32(2.0) * (Aw * up) * D + D * D) * (F32(1.0) + F32(2.0 ** -9))
    kappa = F32(F32(1.0 / math.sqrt(float(B))) * (F32(1.0) + F32(3 * 2.0 ** -23)) * F32(1.4150))   # rsqrtf's unsafe end
    zeta = (F32(7.1) * uu * Aw + yw) * up + F32(2.0 ** -60)
    forced = bool(max(Ax, Ay, Aw, Gx, Gy, Gz) > big)
    ks = F32(kappa * s)
    rows = np.array([F32(P[k] * ks) for k in range(8)] + [F32(P[k] * kappa) for k in range(8, 12)], dtype=F32)
    kz = F32(kappa * zeta)
    gfac = F32(F32(1.99) / F32(kz * kz))
    return rows, gfac, forced


def p_margin(rows, gfac, Xc, nu_s, nv_s):
    x = fma32(rows[0], Xc[0], fma32(rows[1], Xc[1], fma32(rows[2], Xc[2], rows[3])))
    y = fma32(rows[4], Xc[0], fma32(rows[5], Xc[1], fma32(rows[6], Xc[2], rows[7])))
    z = fma32(rows[8], Xc[0], fma32(rows[9], Xc[1], fma32(rows[10], Xc[2], rows[11])))
    a, b = fma32(z, nu_s, x), fma32(z, nv_s, y)
    t = rn32(fr(z) * fr(z))
    mm = fma32(a, a, fma32(b, b, -t))
    g = rn32(fr(t) * fr(gfac))
    return min(mm, g)          # the depth guard: fminf(margin, z'^2 * 1.99 / (kappa zeta)^2)


def cv_pnp_error(m, K4, X32, u32, v32):
    """cv::projectPoints (fp64, un-fused) + the fp32 error, as p_inlier_exact (csrc/score_p.cuh)."""
    fx, fy, cx, cy = (np.float64(v) for v in K4)
    X, Y, Z = (np.float64(X32[:, j]) for j in range(3))
    with np.errstate(all="ignore"):
        x = ((m[0] * X + m[1] * Y) + m[2] * Z) + m[9]
        y = ((m[3] * X + m[4] * Y) + m[5] * Z) + m[10]
        z = ((m[6] * X + m[7] * Y) + m[8] * Z) + m[11]
        z = np.where(z != 0, 1.0 / z, 1.0)
        x = x * z
        y = y * z
        pu, pv = (x * fx + cx).astype(F32), (y * fy + cy).astype(F32)
        dx, dy = u32 - pu, v32 - pv
        return dx * dx + dy * dy


def _rodrigues(v):
    th = np.linalg.norm(v)
    if th < 1e-12:
        return np.eye(3)
    k = v / th
    Kx = np.array([[0, -k[2], k[1]], [k[2], 0, -k[0]], [-k[1], k[0], 0]])
    return np.eye(3) + np.sin(th) * Kx + (1 - np.cos(th)) * Kx @ Kx


def test_pnp_decided_margins_agree_with_opencv():
    """Scenes: the reference's UTM-scale landmarks seen from ~800 m (un-centred coordinates of 7e5 / 2.9e6 m), a desk-scale scene
    (coordinates of order 1), and cameras whose principal plane cuts through the points (depths around zero: the depth guard).
    Pixels are placed so that the squared error is within 1e-8 .. 1e-2 (relative) of the threshold."""
    rng = np.random.default_rng(2027)
    K4 = (4047.87, 2184.27, 982.666819, 697.950868)
    scenes = []
    lo, hi = np.array([738950.0, 2888500.0, 690.0]), np.array([739350.0, 2889050.0, 730.0])
    cam = np.array([739424.6, 2888281.18, 770.0])
    zdir = 0.5 * (lo + hi) - cam
    zdir /= np.linalg.norm(zdir)
    xdir = np.cross(zdir, [0, 0, 1.0]); xdir /= np.linalg.norm(xdir)
    R0 = np.stack([xdir, np.cross(zdir, xdir), zdir])
    scenes.append((lo, hi, R0, cam, 0.0))
    scenes.append((np.array([-0.5, -0.5, 2.0]), np.array([0.5, 0.5, 4.0]), np.eye(3), np.zeros(3), 0.0))
    scenes.append((np.array([-1.0, -1.0, -0.02]), np.array([1.0, 1.0, 0.02]), np.eye(3), np.zeros(3), 1.0))    # depths around zero
    checked = decided = wrong_sign = 0
    worst = 0.0
    n_pts = 40
    for (lo, hi, R0, cam0, plane) in scenes:
        for thr in (F32(64.0), F32(900.0), F32(1.0)):
            for s_bias in (0, 1, -1):
                s = F32(F32(1.0 / math.sqrt(float(thr))) * (F32(1.0) + F32(s_bias * 2.0 ** -22)))
                for h in range(14):
                    R = _rodrigues(rng.normal(0, 10.0 ** rng.uniform(-5, -1), 3)) @ R0
                    cam = cam0 + rng.normal(0, 10.0 ** rng.uniform(-3, 0.5), 3) * (0.0 if plane else 1.0)
                    m = np.concatenate([R.ravel(), -R @ cam])
                    Pw = rng.uniform(lo, hi, (n_pts, 3)).astype(F32)
                    c = Pw.astype(np.float64).mean(axis=0)
                    X = Pw.astype(np.float64)
                    cc = X @ R.T + m[9:]
                    with np.errstate(all="ignore"):
                        pu = K4[0] * cc[:, 0] / cc[:, 2] + K4[2]
                        pv = K4[1] * cc[:, 1] / cc[:, 2] + K4[3]
                    ang = rng.uniform(0, 2 * np.pi, n_pts)
                    rel = 10.0 ** rng.uniform(-8, -2, n_pts) * rng.choice([-1.0, 1.0], n_pts)
                    rad = math.sqrt(float(thr)) * (1.0 + rel)
                    u32 = (pu + rad * np.cos(ang)).astype(F32)
                    v32 = (pv + rad * np.sin(ang)).astype(F32)
                    ok = np.isfinite(u32) & np.isfinite(v32) & (np.abs(u32) < 1e7) & (np.abs(v32) < 1e7)
                    if ok.sum() < 6:
                        continue
                    Pw, u32, v32 = Pw[ok], u32[ok], v32[ok]
                    Xc = (Pw.astype(np.float64) - c).astype(F32)
                    Xm = np.abs(Xc).max(axis=0).astype(F32)
                    rows, gfac, forced = p_filter_const(m, K4, c, Xm, np.abs(u32).max(), np.abs(v32).max(), s)
                    if forced or not np.all(np.isfinite(rows)):
                        continue
                    e = cv_pnp_error(m, K4, Pw, u32, v32)
                    for i in range(len(Pw)):
                        mg = p_margin(rows, gfac, Xc[i], F32(-u32[i] * s), F32(-v32[i] * s))
                        inl_cv = bool(e[i] <= thr)
                        checked += 1
                        if (mg < 0) != inl_cv:
                            wrong_sign += 1
                            worst = max(worst, abs(float(mg)) / 2.0)
                        if abs(float(mg)) >= 2.0:
                            decided += 1
                            assert (mg < 0) == inl_cv, (m, Pw[i], u32[i], v32[i], float(mg), float(e[i]), float(thr))
    assert checked > 8000 and decided > checked // 20 and decided < checked
    assert wrong_sign > 20
    assert worst < 1.0
    print(f"pnp: checked {checked}, decided {decided}, margin sign != cv decision {wrong_sign}, worst |margin|/band {worst:.3f}")
