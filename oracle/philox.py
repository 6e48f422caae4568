"""TEST INFRASTRUCTURE ONLY (see oracle/__init__.py) — numpy restatement of Philox4x32-10 and of the index contract of
the device negative sampler (hassaku_b200/csrc/hsk_sampler.cu).

Philox4x32-10 is the counter-based generator of Salmon et al., "Parallel random numbers: as easy as 1, 2, 3" (SC'11),
as published in Random123 (philox.h: multipliers 0xD2511F53 / 0xCD9E8D57, Weyl key increments 0x9E3779B9 / 0xBB67AE85,
10 rounds).  It is pinned by the Random123 known-answer vectors (kat_vectors) in tests/test_oracle_sampler.py.  The
reference itself samples with numpy's global Mersenne-Twister in loader worker processes (data/dataloader.py:56-57) —
not reproducible across worker counts — so what is pinned against the reference is the SEMANTICS
(data/dataloader.py:110-124: no train item of the user among its negatives; redraw-until-clean; distinct negatives in
the common numpy path), while the index stream is this spec.
"""
import numpy as np

M0, M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
W0, W1 = 0x9E3779B9, 0xBB67AE85
MASK = np.uint64(0xFFFFFFFF)


def philox4x32_10(ctr, key):
    """ctr: (..., 4) uint32, key: (..., 2) uint32 -> (..., 4) uint32."""
    c = [np.asarray(ctr[..., i], dtype=np.uint64) for i in range(4)]
    k0 = np.asarray(key[..., 0], dtype=np.uint64)
    k1 = np.asarray(key[..., 1], dtype=np.uint64)
    for _ in range(10):
        p0 = M0 * c[0]
        p1 = M1 * c[2]
        hi0, lo0 = p0 >> np.uint64(32), p0 & MASK
        hi1, lo1 = p1 >> np.uint64(32), p1 & MASK
        c = [hi1 ^ c[1] ^ k0, lo1, hi0 ^ c[3] ^ k1, lo0]
        k0 = (k0 + np.uint64(W0)) & MASK
        k1 = (k1 + np.uint64(W1)) & MASK
    return np.stack(c, axis=-1).astype(np.uint32)


def _draw(b, j, q, step_lo, k0, k1, n_items):
    """Next uniform item of slot (b, j) (Lemire multiply-shift with rejection); returns (item, new q)."""
    thresh = ((1 << 32) - n_items) % n_items
    while True:
        w = philox4x32_10(np.array([b, j, q >> 2, step_lo], dtype=np.uint32), np.array([k0, k1], dtype=np.uint32))
        r = int(w[q & 3])
        q += 1
        m = r * n_items
        if (m & 0xFFFFFFFF) >= thresh:
            return m >> 32, q


def popularity_cdf_u64(pop_distribution, squashing_factor=1.0):
    """data/dataloader.py:59-64: p = pop^alpha / sum -> 64-bit fixed-point CDF (the array both the kernel and this oracle
    consume).  Computed in Python integers from the float64 probabilities, so it is deterministic."""
    p = np.power(np.asarray(pop_distribution, dtype=np.float64), squashing_factor)
    p = p / p.sum()
    c = np.cumsum(p)
    c = c / c[-1]
    out = np.array([min(int(x * 18446744073709551616.0), 18446744073709551615) for x in c], dtype=np.uint64)
    out[-1] = np.uint64(18446744073709551615)
    return out


def _draw_cdf(b, j, q, step_lo, k0, k1, cdf):
    words = []
    for _ in range(2):
        w = philox4x32_10(np.array([b, j, q >> 2, step_lo], dtype=np.uint32), np.array([k0, k1], dtype=np.uint32))
        words.append(int(w[q & 3]))
        q += 1
    r = (words[0] << 32) | words[1]
    i = int(np.searchsorted(cdf, np.uint64(r), side='right'))
    return min(i, len(cdf) - 1), q


def sample_negatives(u_idx, n_neg, n_items, indptr, indices, seed, step, distinct_in_row=True, max_rounds=256, pop_cdf=None):
    """The device sampler's contract, row by row.  Returns int64 [B, n_neg]."""
    k0 = seed & 0xFFFFFFFF
    k1 = ((seed >> 32) ^ (step >> 32)) & 0xFFFFFFFF
    step_lo = step & 0xFFFFFFFF
    B = len(u_idx)
    out = np.zeros((B, n_neg), dtype=np.int64)
    for b in range(B):
        row = indices[indptr[u_idx[b]]:indptr[u_idx[b] + 1]]
        val = np.zeros(n_neg, dtype=np.int64)
        q = np.zeros(n_neg, dtype=np.int64)
        flagged = np.ones(n_neg, dtype=bool)
        for _ in range(max_rounds):
            drawn = flagged.copy()
            for j in np.nonzero(flagged)[0]:
                if pop_cdf is None:
                    val[j], q[j] = _draw(b, int(j), int(q[j]), step_lo, k0, k1, n_items)
                else:
                    val[j], q[j] = _draw_cdf(b, int(j), int(q[j]), step_lo, k0, k1, pop_cdf)
            flagged = np.zeros(n_neg, dtype=bool)
            flagged[drawn] = np.isin(val[drawn], row)
            if distinct_in_row:
                for j in range(n_neg):
                    if not flagged[j] and (val[j + 1:] == val[j]).any():
                        flagged[j] = True
            if not flagged.any():
                break
        out[b] = val
    return out
