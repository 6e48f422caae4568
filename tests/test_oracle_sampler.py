"""CPU: pins the numpy Philox4x32-10 restatement (oracle/philox.py) with the Random123 known-answer vectors and checks
the sampler contract's semantics against the reference loader's (tests/golden/loader_batches.npz)."""
import numpy as np

from hsk_testutil import load_golden
from oracle import philox as P


def test_philox4x32_10_random123_known_answers():
    kat = [((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
           ((0xffffffff,) * 4, (0xffffffff,) * 2, (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
           ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0),
            (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1))]
    for c, k, e in kat:
        o = P.philox4x32_10(np.array(c, dtype=np.uint32), np.array(k, dtype=np.uint32))
        assert tuple(int(x) for x in o) == e


def test_sampler_contract_semantics_match_reference_loader():
    g = load_golden('loader_batches')
    indptr, indices = g['train_indptr'], g['train_indices']
    u = g['b0/u_idxs']
    neg = P.sample_negatives(u, 20, 200, indptr, indices, seed=64, step=3, distinct_in_row=True)
    ref_neg = g['b0/i_idxs'][:, 1:]
    assert neg.shape == ref_neg.shape and neg.dtype == ref_neg.dtype
    for r in range(len(u)):
        row = indices[indptr[u[r]]:indptr[u[r] + 1]]
        assert not np.isin(neg[r], row).any() and not np.isin(ref_neg[r], row).any()
        assert len(set(neg[r])) == 20            # distinct, like the reference's common numpy path ...
    assert (neg >= 0).all() and (neg < 200).all()
    # deterministic in (seed, step); different steps give different draws
    assert (P.sample_negatives(u, 20, 200, indptr, indices, 64, 3) == neg).all()
    assert (P.sample_negatives(u, 20, 200, indptr, indices, 64, 4) != neg).any()
    # without the distinct rule duplicates may survive, train items never do
    neg2 = P.sample_negatives(u, 60, 200, indptr, indices, 1, 0, distinct_in_row=False)
    assert any(len(set(r)) < 60 for r in neg2)


def test_sampler_contract_is_uniform_over_allowed_items():
    indptr = np.array([0, 3], dtype=np.int64)
    indices = np.array([1, 5, 7], dtype=np.int32)
    draws = np.concatenate([P.sample_negatives(np.zeros(64, dtype=np.int64), 8, 16, indptr, indices, 5, s,
                                               distinct_in_row=False).ravel() for s in range(12)])
    counts = np.bincount(draws, minlength=16)
    assert counts[[1, 5, 7]].sum() == 0
    allowed = np.delete(counts, [1, 5, 7])
    exp = allowed.sum() / 13
    chi2 = ((allowed - exp) ** 2 / exp).sum()
    assert chi2 < 34.5  # chi-square, 12 dof, p = 0.0005
