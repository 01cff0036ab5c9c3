"""ORACLE — TEST INFRASTRUCTURE ONLY.  Philox4x32-10 and the throughput-mode minimal-sample generator, restated in NumPy.

The reference has no counter-based sampler (OpenCV draws its samples from cv::RNG, restated in cv_ransac_oracle.c);
BASELINE.json's north star asks for one ("a Philox counter-based sample generator that can also replay the reference's
exact sample indices").  Philox4x32-10 is the published algorithm of Salmon, Moraes, Dror, Shaw, "Parallel random numbers:
as easy as 1, 2, 3" (SC'11; Random123 library, philox.h): ten rounds of
    (c0, c1, c2, c3) <- (hi(M1 c2) ^ c1 ^ k0, lo(M1 c2), hi(M0 c0) ^ c3 ^ k1, lo(M0 c0)),  k0 += W0, k1 += W1
with M0 = 0xD2511F53, M1 = 0xCD9E8D57, W0 = 0x9E3779B9, W1 = 0xBB67AE85.  KNOWN_ANSWERS are the three philox4x32-10 lines
of Random123's kat_vectors file; tests/test_host_logic.py checks this restatement against them, and the GPU sampler
(csrc/sampler.cuh: philox4x32_10, distinct4; csrc/pipeline_h.cuh: k_philox_sample_solve_h) against this restatement."""
import numpy as np

M0, M1, W0, W1 = 0xD2511F53, 0xCD9E8D57, 0x9E3779B9, 0xBB67AE85
MAX_ATTEMPTS = 16   # PHILOX_MAX_ATTEMPTS, csrc/sampler.cuh

# (counter, key, expected output) — Random123 kat_vectors, "philox4x32 10"
KNOWN_ANSWERS = [
    ((0x00000000, 0x00000000, 0x00000000, 0x00000000), (0x00000000, 0x00000000),
     (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
    ((0xffffffff, 0xffffffff, 0xffffffff, 0xffffffff), (0xffffffff, 0xffffffff),
     (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
    ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0),
     (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1)),
]


def philox4x32_10(c0, c1, c2, c3, k0, k1):
    """Vectorised over arrays of counters; returns four uint32 arrays."""
    c = [np.asarray(x, dtype=np.uint64) & 0xFFFFFFFF for x in np.broadcast_arrays(c0, c1, c2, c3)]
    k0, k1 = int(k0) & 0xFFFFFFFF, int(k1) & 0xFFFFFFFF
    for _ in range(10):
        p0, p1 = np.uint64(M0) * c[0], np.uint64(M1) * c[2]
        c = [(p1 >> np.uint64(32)) ^ c[1] ^ np.uint64(k0), p1 & np.uint64(0xFFFFFFFF),
             (p0 >> np.uint64(32)) ^ c[3] ^ np.uint64(k1), p0 & np.uint64(0xFFFFFFFF)]
        k0, k1 = (k0 + W0) & 0xFFFFFFFF, (k1 + W1) & 0xFFFFFFFF
    return [x.astype(np.uint32) for x in c]


def distinct(words, n):
    """k distinct indices in [0, n) from k uniform 32-bit words, no rejection: the t-th draw picks the j-th not yet chosen
    index, j = floor(word_t * (n - t) / 2^32)."""
    chosen, out = [], []
    for t, w in enumerate(words):
        j = (int(w) * (n - t)) >> 32
        for c in sorted(chosen):
            if j >= c:
                j += 1
        chosen.append(j)
        out.append(j)
    return out


def sample_h(src_f32, dst_f32, seed, hyp_begin, n_hyp, check_subset, q=0):
    """The 4-point samples of hypothesis ids [hyp_begin, hyp_begin + n_hyp) of problem q: attempt a of hypothesis g uses
    philox(counter = (g_lo, g_hi, a, q), key = seed); the first attempt whose subset passes OpenCV's checkSubset
    (check_subset(src4, dst4) -> bool, the oracle's) is the sample; all -1 after MAX_ATTEMPTS failures."""
    n = len(src_f32)
    out = np.full((n_hyp, 4), -1, dtype=np.int32)
    for g in range(n_hyp):
        gid = hyp_begin + g
        for a in range(MAX_ATTEMPTS):
            w = philox4x32_10(gid & 0xFFFFFFFF, gid >> 32, a, q, seed & 0xFFFFFFFF, seed >> 32)
            idx = distinct([int(x) for x in w], n)
            if check_subset(src_f32[idx], dst_f32[idx]):
                out[g] = idx
                break
    return out
