#!/usr/bin/env python
"""Generates tests/golden/cv2_lm_blocks.json — run in the BUILD CONTAINER only (needs cv2 4.13.0).

Bit-level known answers for the building blocks of cv::LMSolver as cv2.findHomography runs it after its RANSAC stage
(refit + 10 Levenberg-Marquardt iterations; the reference's call is /root/reference/main_v1.py:312), taken from the
functions the cv2 4.13.0 binary exposes:
  cv2.mulTransposed(J, True)                 A = J^T J
  cv2.gemm(J, r, 1, None, 0, GEMM_1_T)       v = J^T r
  cv2.gemm(A, d, -1, v, 2)                   2 v - A d
  cv2.norm(r, NORM_L2SQR)                    |r|^2   (AVX2-dispatched: fused multiply-adds in a fixed lane order)
  cv2.solve(A, v, DECOMP_EIG), cv2.invert(A, DECOMP_EIG)
and for the whole refinement: cv2.findHomography(src, dst, 0) (= runKernel + the LM, no RANSAC) on seeded problems of
5 ... 39 points, ill-conditioned ones included.  tests/test_oracle_golden.py replays all of it against the oracle
WITHOUT cv2 and demands equality of every float64 bit.  Floats are written as hex strings (float.hex)."""
import json
import os

import cv2
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))


def hx(a):
    return [float(x).hex() for x in np.asarray(a, dtype=np.float64).ravel()]


def main():
    rng = np.random.default_rng(413)
    g = {"cv2_version": cv2.__version__, "norm_l2sqr": [], "gemm_atb": [], "gemm_axpby": [], "mul_transposed": [],
         "solve_eig": [], "invert_eig_diag": [], "find_homography_0": []}
    for n in list(range(1, 41)) + [47, 48, 49, 63, 64, 65, 100, 129]:
        r = rng.standard_normal((n, 1)) * 10 ** rng.uniform(-3, 3, (n, 1))
        g["norm_l2sqr"].append(dict(r=hx(r), out=float(cv2.norm(r, cv2.NORM_L2SQR)).hex()))
    for rows in (10, 12, 14, 16, 18, 22, 24, 40, 56):
        J = rng.standard_normal((rows, 9)) * 10 ** rng.uniform(-3, 3, (1, 9))
        r = rng.standard_normal((rows, 1))
        g["gemm_atb"].append(dict(rows=rows, J=hx(J), r=hx(r), out=hx(cv2.gemm(J, r, 1, None, 0, flags=cv2.GEMM_1_T))))
        g["mul_transposed"].append(dict(rows=rows, J=hx(J), out=hx(cv2.mulTransposed(J, True))))
    for _ in range(12):
        J = rng.standard_normal((14, 9)) * 10 ** rng.uniform(-2, 2, (1, 9))
        A = cv2.mulTransposed(J, True)
        d, v = rng.standard_normal((9, 1)), rng.standard_normal((9, 1))
        g["gemm_axpby"].append(dict(A=hx(A), d=hx(d), c=hx(v), out=hx(cv2.gemm(A, d, -1, v, 2))))
        Ap = A.copy()
        lam = float(rng.choice([0.0, 1.0, 0.37]))
        for i in range(9):
            Ap[i, i] += lam * A[i, i]
        g["solve_eig"].append(dict(A=hx(Ap), b=hx(v), out=hx(cv2.solve(Ap, v, flags=cv2.DECOMP_EIG)[1])))
        g["invert_eig_diag"].append(dict(A=hx(A), out=hx(np.diag(cv2.invert(A, flags=cv2.DECOMP_EIG)[1]))))
    for t in range(160):
        n = int(rng.integers(5, 40))
        src = rng.uniform(-1, 1, (n, 2)) * 10 ** rng.uniform(-1, 3)
        Ht = np.eye(3) + rng.standard_normal((3, 3)) * 0.1
        Ht[2, :2] *= 1e-3
        q = np.c_[src, np.ones(n)] @ Ht.T
        dst = q[:, :2] / q[:, 2:] + rng.standard_normal((n, 2)) * 10 ** rng.uniform(-3, 1)
        src = src.astype(np.float32).astype(np.float64)
        dst = dst.astype(np.float32).astype(np.float64)
        H, _ = cv2.findHomography(src, dst, 0)
        if H is None:
            continue
        g["find_homography_0"].append(dict(src=hx(src), dst=hx(dst), H=hx(H)))
    with open(os.path.join(HERE, "cv2_lm_blocks.json"), "w") as f:
        json.dump(g, f)
    print({k: (len(v) if isinstance(v, list) else v) for k, v in g.items()})


if __name__ == "__main__":
    main()
