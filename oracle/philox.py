"""TEST INFRASTRUCTURE ONLY -- CPU oracle, never imported by the product path.

Philox4x32-10 counter-based generator (Salmon, Moraes, Dror, Shaw, SC'11; the
Random123 "philox4x32_10"), restated in numpy from the published algorithm, and the
sampler's addressing convention: key = 64-bit seed, counter =
(cell_lo, cell_hi, sub, iter << 8 | purpose).

The reference (R) draws from the global Mersenne-Twister stream through
stats::r* (e.g. /root/reference/R/sample_params.R:263 `rmultinom`,
R/sample_Pn.R:116-118 `rgamma`); a counter-based stream is what makes the GPU
sampler independent of thread/shard layout, so parity with the reference is
distributional and parity oracle<->GPU is draw-for-draw.

Pinned against the Random123 known-answer vectors in tests/test_oracle_philox.py.
"""
import numpy as np

M0 = np.uint64(0xD2511F53)
M1 = np.uint64(0xCD9E8D57)
W0 = np.uint32(0x9E3779B9)
W1 = np.uint32(0xBB67AE85)
MASK32 = np.uint64(0xFFFFFFFF)

# draw purposes (must match bayesnmf_b200/csrc/bnmf_rng.cuh)
PUR_Z, PUR_P, PUR_E = 0, 1, 2
PUR_HYP_P1, PUR_HYP_P2, PUR_HYP_E1, PUR_HYP_E2 = 3, 4, 5, 6
PUR_MH_P, PUR_MH_E, PUR_A, PUR_R, PUR_SIGMASQ = 7, 8, 9, 10, 11


def philox4x32_10(c0, c1, c2, c3, k0, k1):
    """All arguments broadcastable uint32 arrays; returns 4 uint32 arrays."""
    c0, c1, c2, c3 = np.broadcast_arrays(*(np.asarray(x, dtype=np.uint32) for x in (c0, c1, c2, c3)))
    c0, c1, c2, c3 = c0.copy(), c1.copy(), c2.copy(), c3.copy()
    k0 = np.uint32(k0)
    k1 = np.uint32(k1)
    with np.errstate(over="ignore"):
        for _ in range(10):
            p0 = M0 * c0.astype(np.uint64)
            p1 = M1 * c2.astype(np.uint64)
            hi0 = (p0 >> np.uint64(32)).astype(np.uint32)
            lo0 = (p0 & MASK32).astype(np.uint32)
            hi1 = (p1 >> np.uint64(32)).astype(np.uint32)
            lo1 = (p1 & MASK32).astype(np.uint32)
            n0 = hi1 ^ c1 ^ k0
            n2 = hi0 ^ c3 ^ k1
            c0, c1, c2, c3 = n0, lo1, n2, lo0
            k0 = np.uint32((int(k0) + int(W0)) & 0xFFFFFFFF)
            k1 = np.uint32((int(k1) + int(W1)) & 0xFFFFFFFF)
    return c0, c1, c2, c3


def words(seed, it, purpose, cell, sub):
    """Philox block `sub` of stream (it, purpose, cell).  cell/sub broadcastable ints."""
    cell = np.asarray(cell, dtype=np.uint64)
    sub = np.asarray(sub, dtype=np.uint64)
    seed = int(seed) & 0xFFFFFFFFFFFFFFFF
    c3 = np.uint32(((int(it) << 8) | int(purpose)) & 0xFFFFFFFF)
    return philox4x32_10(
        (cell & MASK32).astype(np.uint32),
        (cell >> np.uint64(32)).astype(np.uint32),
        (sub & MASK32).astype(np.uint32),
        c3,
        seed & 0xFFFFFFFF,
        seed >> 32,
    )


def u01(w, bits=32):
    """Open-interval uniform from a 32-bit word.  bits=24 emulates the float path."""
    w = np.asarray(w, dtype=np.uint32)
    if bits == 32:
        return (w.astype(np.float64) + 0.5) * 2.0 ** -32
    return ((w >> np.uint32(8)).astype(np.float64) + 0.5) * 2.0 ** -24
