"""Cycle breakdown of k_finalize_h by section (development probe).

`python tools/prof_finalize_sections.py build` (here: nvcc cross-compiles) builds a copy of the library with
-DB2R_FIN_PROFILE into code-reproduction-ransac_b200/libransac_b200_prof.so; `python tools/prof_finalize_sections.py`
(on the GPU box) loads that copy instead of the product and prints, per problem shape, the clock64() deltas thread 0 of
CTA 0 accumulated in every section of the kernel (FINCLK markers in csrc/pipeline_h.cuh) next to the stage times."""
import ctypes as C
import importlib.util
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "code-reproduction-ransac_b200")
PROF_LIB = os.path.join(PKG, "libransac_b200_prof.so")
SECTIONS = {0: "mask+count", 1: "norm stats", 2: "LtL pass+reduce", 3: "LtL eigvec", 4: "pre-eval0", 5: "loop control", 6: "cholesky step",
            7: "eig step", 8: "gain ratio (thread 0)", 9: "inverse diagonal", 10: "accept + stop test", 11: "eval: point pass",
            12: "eval: reduction", 13: "eval: assemble (thread 0)", 14: "final mask + outputs"}


def _build_mod():
    spec = importlib.util.spec_from_file_location("b2r_build", os.path.join(PKG, "_build.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def build():
    b = _build_mod()
    objdir = os.path.join(PKG, "build", "prof")
    os.makedirs(objdir, exist_ok=True)
    objs, procs = [], []
    for tu in b.TRANSLATION_UNITS:
        obj = os.path.join(objdir, tu.replace(".cu", ".o"))
        procs.append(subprocess.Popen(["nvcc"] + b.NVCC_FLAGS + ["-DB2R_FIN_PROFILE", "-c", "-o", obj, os.path.join(b.CSRC, tu)]))
        objs.append(obj)
    for p in procs:
        if p.wait() != 0:
            raise SystemExit("nvcc failed")
    subprocess.check_call(["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o", PROF_LIB] + objs)
    print(PROF_LIB)


def main():
    sys.path.insert(0, ROOT)
    import numpy as np
    import ransac_b200
    from ransac_b200 import _build, _lib, synth
    _build.LIB_PATH = PROF_LIB          # the instrumented copy, not the product
    _build.needs_build = lambda: False
    ctx = ransac_b200.Context(0)
    lib = _lib.load()
    fn = lib.b2r_debug_fin_clocks
    fn.restype = C.c_int
    fn.argtypes = [C.POINTER(C.c_ulonglong), C.c_int]
    buf = (C.c_ulonglong * 32)()
    out = {}
    shapes = [(1000, 10_000, 0.3), (20_000, 10_000, 0.3), (100_000, 100_000, 0.5)]
    reps = 5
    for n, H, outl in shapes:
        src, dst, _ = synth.homography_set(n, outl, np.random.default_rng(1899))
        prob = ctx.upload(src[None], dst)
        par = ransac_b200.make_params(3.0, H, sampler=ransac_b200.SAMPLER_PHILOX, seed=3, arith=ransac_b200.ARITH_FAST,
                                      solver=ransac_b200.SOLVER_FAST)
        for _ in range(3):
            prob.run(par); prob.fetch()
        fn(buf, 1)
        for _ in range(reps):
            prob.run(par); prob.fetch()
        fn(buf, 1)
        _, _, info = prob.fetch()
        cyc = {SECTIONS[i]: int(buf[i]) // reps for i in SECTIONS if buf[i]}
        tot = sum(cyc.values())
        out[f"n{n}"] = {"lm_iters": int(info[0]["lm_iters"]), "finalize_ms": round(prob.stage_ms()["finalize"], 4), "cycles_total": tot,
                        "us_at_1965MHz": round(tot / 1965.0, 1), "cycles": cyc}
        prob.free()
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "build":
        build()
    else:
        main()
