"""CPU: the C ABI loads and exports every symbol the header declares; host-side logic (geodesy, the reference's
score formulas, hypothesis sharding, the gloo path of the multi-rank reduce).  No compute calls without a GPU."""
import os
import re
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    text = open(os.path.join(ROOT, "include", "ransac_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(b2r_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    import __graft_entry__ as g
    g.build()
    from ransac_b200 import _lib
    L = _lib.load()
    declared = header_symbols()
    assert len(declared) >= 25
    assert sorted(_lib.SIGNATURES) == declared          # the ctypes table and the header agree
    for name in declared:
        assert getattr(L, name) is not None
    out = subprocess.run(["nm", "-D", "--defined-only", _lib.lib_path()], capture_output=True, text=True).stdout
    exported = sorted(set(re.findall(r" T (b2r_[a-z0-9_]+)", out)))
    assert exported == declared                          # and nothing else is exported under that prefix


def test_scoring_kernels_are_the_instruction_mix_design_md_describes():
    """SASS of the built library (cuobjdump; no GPU needed): the fast 3x3 scoring kernel is packed fp32x2 arithmetic with
    no reciprocal and sign-bit counting (11 FFMA2/FMUL2 per pair of hypotheses and point, LEA.HI, no MUFU in the loop),
    it is fed by a bulk TMA copy, the exact kernel never contracts a product into its sum (every FFMA2 of its batch adds
    RZ or belongs to the Newton step), and no scoring kernel touches the tensor pipe or spills."""
    import shutil
    if shutil.which("cuobjdump") is None:
        pytest.skip("cuobjdump not on PATH")
    import __graft_entry__ as g
    g.build()
    from ransac_b200 import _lib
    sass = subprocess.run(["cuobjdump", "-sass", _lib.lib_path()], capture_output=True, text=True).stdout
    funcs = {}
    name = None
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            name = m.group(1)
            funcs[name] = []
        elif name and re.search(r"/\*[0-9a-f]{4}\*/", line):
            funcs[name].append(line)
    fast = next(v for k, v in funcs.items() if "k3_score_hILi2ELb0" in k)
    exact = next(v for k, v in funcs.items() if "k3_score_hILi2ELb1" in k)
    pfast = next(v for k, v in funcs.items() if "k3_score_p_fastILi2" in k)

    def count(lines, op):
        return sum(1 for l in lines if re.search(r"\b" + op + r"\b", l))

    def main_loop(lines):   # from the first broadcast LDS.128 to the backward branch that closes the unrolled loop
        start = next(i for i, l in enumerate(lines) if "LDS.128" in l)
        end = next(i for i in range(start, len(lines)) if re.search(r"\bBRA\b", lines[i]))
        return lines[start:end]

    loop = main_loop(fast)
    n_lds = count(loop, r"LDS\.128")
    assert n_lds == 4                                                    # unrolled by 4 points
    assert count(loop, "FFMA2") + count(loop, "FMUL2") == 11 * 2 * n_lds  # 11 packed ops x 2 pairs per point
    assert count(loop, r"MUFU\.RCP") == 0 and count(loop, "FSETP") == 0
    assert count(loop, r"LEA\.HI") == 4 * n_lds                          # one per evaluation
    for body in (fast, exact, pfast):
        assert any("UBLKCP" in l for l in body)                           # 1-D bulk TMA copy of the point tile
        assert not any(re.search(r"\b(HMMA|IMMA|DMMA|UTCHMMA|UTCMMA|STL|LDL)\b", l) for l in body)
    ploop = main_loop(pfast)
    assert count(ploop, r"MUFU\.RCP") == 0 and count(ploop, "FFMA2") + count(ploop, "FMUL2") == 14 * 2 * count(ploop, r"LDS\.128")
    eloop = main_loop(exact)
    n_pts = count(eloop, r"LDS\.128")
    assert n_pts == 8
    assert count(eloop, "FADD2") == 9 * 2 * n_pts                         # every sum is its own instruction
    ffma2 = [l for l in eloop if re.search(r"\bFFMA2\b", l)]
    assert len(ffma2) == 12 * 2 * n_pts                                   # 10 products (+ RZ) and the 2 FMAs of the Newton step
    assert sum(1 for l in ffma2 if re.search(r", RZ(\.F32)? ;", l) or ", RZ ;" in l) >= 10 * 2 * n_pts
    # the filtered exact kernels (score_h_filt.cuh, score_p_filt.cuh): the fast kernels' packed arithmetic, ONE funnel shift per
    # evaluation that files sign and band bit (no per-evaluation compare, min or count), no reciprocal, no fp64 and no local
    # memory in the loop; the 3x4 kernel adds the depth guard (one FMUL2 per pair, one FMNMX per evaluation)
    hfilt = next(v for k, v in funcs.items() if "k3_score_h_filtILi2" in k)
    pfilt = next(v for k, v in funcs.items() if "k3_score_p_filtILi2" in k)

    def unrolled_loop(lines, per_point):   # the longest run of instructions between two branches that holds packed arithmetic
        best, cur = [], []
        for l in lines:
            if re.search(r"\b(BRA|EXIT|RET|CALL)\b", l):
                if count(cur, "FFMA2") > count(best, "FFMA2"):
                    best = cur
                cur = []
            else:
                cur.append(l)
        assert count(best, "FFMA2") >= per_point
        return best

    floop = unrolled_loop(hfilt, 20)
    n_pts = count(floop, r"LDS\.128")
    assert n_pts >= 4
    assert count(floop, "FFMA2") + count(floop, "FMUL2") == 11 * 2 * n_pts
    assert count(floop, r"SHF\.L\.W\.U32\.HI") == 4 * n_pts             # one per evaluation: (sg << 2) | (margin >> 30)
    for op in (r"MUFU\.\w+", "FSETP", "FMNMX3?", r"LEA\.HI", "POPC", "DFMA", "DADD", "DMUL", "STL", "LDL", r"ATOMS?\.\w+", "BAR"):
        assert count(floop, op) == 0, op
    ploop2 = unrolled_loop(pfilt, 26)
    n_pts = count(ploop2, r"LDS\.128")
    assert n_pts >= 2
    assert count(ploop2, "FFMA2") + count(ploop2, "FMUL2") == 15 * 2 * n_pts   # 14 of the margin + the depth guard's product
    assert count(ploop2, r"SHF\.L\.W\.U32\.HI") == 4 * n_pts and count(ploop2, "FMNMX") == 4 * n_pts
    for op in (r"MUFU\.\w+", "FSETP", r"LEA\.HI", "DFMA", "DADD", "DMUL", "STL", "LDL", r"ATOMS?\.\w+", "BAR"):
        assert count(ploop2, op) == 0, op
    for body in (hfilt, pfilt):
        assert any("UBLKCP" in l for l in body)
        assert not any(re.search(r"\b(HMMA|IMMA|DMMA|UTCHMMA|UTCMMA)\b", l) for l in body)


def test_no_cpu_fallback():
    """Without a CUDA device the product path must fail loudly (and it never imports the oracle)."""
    import ransac_b200
    from ransac_b200 import _lib
    if _lib.load().b2r_device_count() > 0:
        pytest.skip("a CUDA device is present")
    with pytest.raises(ransac_b200.RansacB200Error, match="no CPU fallback"):
        ransac_b200.Context(0)
    pkg = os.path.join(ROOT, "code-reproduction-ransac_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh")):
                src = open(os.path.join(dirpath, f), encoding="utf-8").read()
                assert not re.search(r"^\s*(import|from)\s+oracle\b", src, flags=re.M), f


def test_info_records_are_lazy_but_list_like():
    """Batched calls hand back their per-problem info records as a sequence that builds a dict on access; it must still
    behave like the list of dicts the callers index, slice and iterate (pipeline.py, dist.py, the tests)."""
    from ransac_b200 import api, _lib
    arr = (_lib.HInfo * 4)()
    for q in range(4):
        arr[q].status, arr[q].iters_run, arr[q].best_count = (1 if q == 2 else 0), 10 * q, q
        for j in range(4):
            arr[q].sample[j] = q + j
    seq = api._InfoSeq(arr, api._info_dict)
    assert len(seq) == 4 and seq.status.tolist() == [0, 0, 1, 0]
    assert seq[3]["iters_run"] == 30 and seq[1]["sample"] == [1, 2, 3, 4]
    assert [d["best_count"] for d in seq] == [0, 1, 2, 3]
    assert [d["iters_run"] for d in seq[1:3]] == [10, 20]
    parr = (_lib.PInfo * 2)()
    parr[1].status, parr[1].mean_inlier_err = -3, 2.5
    pseq = api._InfoSeq(parr, api._p_info_dict)
    assert pseq.status.tolist() == [0, -3] and pseq[1]["mean_inlier_err"] == 2.5
    assert len(api._InfoSeq((_lib.HInfo * 0)(), api._info_dict)) == 0


def test_utm_transform_check_values():
    from ransac_b200 import geo
    e, n = geo.wgs84_to_utm50n(119.390036, 26.098989)           # testpro-K.py:199
    assert abs(e - 739031.2) < 5e-3 and abs(n - 2888840.39) < 5e-3
    e, n = geo.wgs84_to_utm50n(119.39055048629785, 26.09361027127146)
    assert abs(e - 739093.6175) < 1e-3 and abs(n - 2888245.3439) < 1e-3


def test_candidate_projection_and_score_formulas():
    """candidate_pos2 / _score restate main_v1.py:304-311 and :327-348, :419 — checked against a literal loop."""
    from ransac_b200 import pipeline
    rng = np.random.default_rng(0)
    pos3d = rng.uniform([738950, 2888500, 690], [739350, 2889050, 730], (12, 3))
    cam = np.array([739410.15, 2888321.95, 756.0])
    pixels = rng.uniform(0, 2000, (12, 2))
    pos2 = pipeline.candidate_pos2(pos3d, cam)
    for i in range(12):
        p = pos3d[i] - cam
        p = np.array([p[2], p[1], p[0]])
        p = p / p[2]
        np.testing.assert_array_equal(pos2[i], p[0:2])
    H = np.array([[1500., 80, 600], [-40, -700, 100], [0.05, -0.02, 1]])
    mask = (rng.random(12) < 0.7).astype(np.uint8).reshape(-1, 1)
    M, e1, e2 = pipeline._score(H, mask, pos2, pixels, 75.0)
    Mr = np.linalg.inv(H)
    r1 = r2 = 0.0
    for i in range(12):
        pp2 = np.linalg.inv(Mr) @ np.array([pos2[i, 0], pos2[i, 1], 1.0])
        pp2 = pp2 / pp2[2]
        PP2 = Mr @ np.array([pixels[i, 0], pixels[i, 1], 1.0])
        PP2 = PP2 / PP2[2]
        if mask[i] == 1:
            r1 += np.linalg.norm(pixels[i] - pp2[0:2])
            r2 += np.linalg.norm(pos2[i] - PP2[0:2])
    r2 += np.sum(1 - mask) * 75.0
    assert abs(e1 - r1) < 1e-9 * max(1, r1) and abs(e2 - r2) < 1e-9 * max(1, r2)
    nm = np.array([[1., 0.], [2., 5.], [3., 4.]])
    assert pipeline.best_location(nm) == 2                       # zeros count as 1e6, main_v1.py:865


def test_shard_range_partitions_ids():
    from ransac_b200 import dist
    for total, world in [(100000, 8), (10, 3), (7, 8), (1, 1)]:
        covered = []
        for r in range(world):
            b, c = dist.shard_range(total, r, world)
            covered += list(range(b, b + c))
        assert covered == list(range(total))


WORKER = r"""
import os, sys
sys.path.insert(0, {root!r})
import numpy as np
import torch.distributed as dist
from ransac_b200 import dist as rdist
rank = int(os.environ["RANK"])
dist.init_process_group("gloo", init_method="tcp://127.0.0.1:{port}", rank=rank, world_size=2)
# rank 0 holds the better count for problem 0; equal counts for problem 1 -> the lower hypothesis id wins
keys = [np.array([(50 << 32) | (0xFFFFFFFF - 7), (40 << 32) | (0xFFFFFFFF - 9)], dtype=np.uint64),
        np.array([(48 << 32) | (0xFFFFFFFF - 3), (40 << 32) | (0xFFFFFFFF - 2)], dtype=np.uint64)][rank]
best = rdist.reduce_keys_max(keys)
assert int(best[0]) >> 32 == 50 and 0xFFFFFFFF - (int(best[0]) & 0xFFFFFFFF) == 7, best
assert int(best[1]) >> 32 == 40 and 0xFFFFFFFF - (int(best[1]) & 0xFFFFFFFF) == 2, best
assert rdist.shard_range(10, rank, 2) == ((0, 5) if rank == 0 else (5, 5))
# problem-sharded batched calls: every rank contributes its block of per-problem scores, all get the whole table
table = np.arange(7 * 2, dtype=np.float64).reshape(7, 2) * 1.5
b, c = rdist.shard_range(7, rank, 2)
got = rdist.allgather_rows(table[b:b + c], 7)
assert np.array_equal(got, table), got
dist.destroy_process_group()
print("ok", rank)
"""


def test_two_rank_key_reduce_over_gloo(tmp_path):
    """The N>1 path of dist.reduce_keys_max with world_size 2 on CPU (gloo): MAX of the packed keys = best count,
    lowest id on ties."""
    import socket
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    script = tmp_path / "worker.py"
    script.write_text(WORKER.format(root=ROOT, port=port))
    procs = []
    for r in range(2):
        env = dict(os.environ, RANK=str(r), WORLD_SIZE="2", MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
        procs.append(subprocess.Popen([sys.executable, str(script)], env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True))
    outs = [p.communicate(timeout=180)[0] for p in procs]
    for r, (p, o) in enumerate(zip(procs, outs)):
        assert p.returncode == 0, o
        assert f"ok {r}" in o


def test_philox_known_answers():
    """Philox4x32-10 (north-star kernel 1) against the three known-answer vectors of Random123's kat_vectors; the GPU
    sampler is compared with this restatement in tests/test_gpu_parity_h.py::test_philox_sampler_matches_restatement."""
    from oracle import philox
    for ctr, key, want in philox.KNOWN_ANSWERS:
        got = philox.philox4x32_10(*ctr, *key)
        assert [int(x) for x in got] == list(want)
    # vectorised form == scalar form
    ids = np.arange(5, dtype=np.uint64) + np.uint64(0xFFFFFFFE)
    v = philox.philox4x32_10(ids & np.uint64(0xFFFFFFFF), ids >> np.uint64(32), 3, 0, 7, 0)
    for i, g in enumerate(ids):
        s = philox.philox4x32_10(int(g) & 0xFFFFFFFF, int(g) >> 32, 3, 0, 7, 0)
        assert [int(x[i]) for x in v] == [int(x) for x in s]
    # distinct(): k distinct indices in range, uniform mapping of the words
    rng = np.random.default_rng(0)
    for n in (4, 5, 12, 1000):
        for _ in range(200):
            idx = philox.distinct(rng.integers(0, 2**32, 4).tolist(), n)
            assert len(set(idx)) == 4 and min(idx) >= 0 and max(idx) < n
    assert philox.distinct([0, 0, 0, 0], 4) == [0, 1, 2, 3] and philox.distinct([2**32 - 1] * 4, 4) == [3, 2, 1, 0]
