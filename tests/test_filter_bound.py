"""The error bound of the filtered exact predicate (csrc/score_h_filt.cuh), checked on the CPU in exact arithmetic.

The kernel accepts the sign of its division-free fp32 FMA margin when |margin| >= 2.0 after scaling the hypothesis by
(2.002 / B)^1/2.  Here that margin is re-computed with correctly rounded fp32 FMAs built from Python fractions (no GPU, no
libm), the per-hypothesis scale from the same fp32 operations as k3_filter_const(), and OpenCV's comparison with numpy's
individually rounded float32 operations (HomographyEstimatorCallback::computeError; reference call site
main_v1.py:312).  Pixels are placed so that the squared error lands within 1e-8 .. 1e-2 (relative) of the threshold, on
hypotheses with the cancellation of the reference's pos2 coordinates, with vanishing denominators, and with huge and tiny
coefficients.  Claim under test: decided  =>  (margin < 0) == (err_cv <= thr).  The slack of the bound is reported."""
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
