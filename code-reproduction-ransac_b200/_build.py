"""Build recipe of libransac_b200.so: nvcc, sm_100a only, in-tree (the .so travels with the repo snapshot)."""
import os
import shutil
import subprocess

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG_DIR, "csrc")
LIB_PATH = os.path.join(PKG_DIR, "libransac_b200.so")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo",
    # never contract a*b+c on its own: the bit-exact paths need every product and sum rounded separately;
    # fused operations are written explicitly (fma.rn.f32x2 / __fmaf_rn) where they are wanted
    "-fmad=false",
    "-Xcompiler", "-fPIC,-fvisibility=hidden",
]
TRANSLATION_UNITS = ["api.cu", "api_pnp.cu", "api_geo.cu"]   # homography path, PnP path, DEM ray-march


def _sources():
    out = [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC)) if f.endswith((".cu", ".cuh", ".h"))]
    out.append(os.path.join(PKG_DIR, "..", "include", "ransac_b200.h"))
    return out


def needs_build():
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    return any(os.path.getmtime(s) > t for s in _sources())


def build(force=False, verbose=False):
    """Compile the translation units of csrc/ (in parallel) and link them into libransac_b200.so."""
    if not force and not needs_build():
        return LIB_PATH
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    objdir = os.path.join(PKG_DIR, "build")
    os.makedirs(objdir, exist_ok=True)
    procs = []
    for tu in TRANSLATION_UNITS:
        obj = os.path.join(objdir, tu.replace(".cu", ".o"))
        cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-c", "-o", obj, os.path.join(CSRC, tu)]
        procs.append((tu, obj, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    objs = []
    for tu, obj, pr in procs:
        out, _ = pr.communicate()
        if pr.returncode != 0:
            raise RuntimeError(f"nvcc failed on {tu}:\n" + out)
        if verbose:
            print(out)
        objs.append(obj)
    cmd = [nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o", LIB_PATH] + objs
    res = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc link failed:\n" + res.stdout)
    return LIB_PATH
