"""Seeded synthetic implicit-feedback data of the shapes BASELINE.json names.

Users uniform, items Zipf-like p(i) ∝ (rank+1)^-0.8, de-duplicated (u, i) pairs, per-interaction
uniform r in [0,1) -> 80/10/10 train/val/test (mirrors the ratio split of the reference's
data/data_utils.py:241-277).  `write_csv_dataset` emits the 5-file layout the reference's loaders
read (data/dataset.py:10-23) so the *unmodified* reference can consume the same data.
"""
from __future__ import annotations

import os
from dataclasses import dataclass
from typing import Optional

import numpy as np
from scipy import sparse as sp

SHAPES = {
    # name: (n_users, n_items, n_interactions)
    'ml1m': (6040, 3706, 1_000_000),
    'ml10m': (69_878, 10_677, 10_000_000),
    'lfm2b': (2_000_000, 1_000_000, 200_000_000),
    'tiny': (300, 200, 6_000),
}


@dataclass
class SyntheticInteractions:
    n_users: int
    n_items: int
    train: sp.csr_matrix  # bool/int8 CSR, sorted indices
    val: sp.csr_matrix
    test: sp.csr_matrix
    user_group: Optional[np.ndarray]  # [n_users] int64 or None
    n_user_groups: int


def _csr(u, i, n_users, n_items):
    m = sp.csr_matrix((np.ones(len(u), dtype=np.int8), (u, i)), shape=(n_users, n_items))
    m.sort_indices()
    return m


def make_interactions(n_users: int, n_items: int, n_interactions: int, seed: int = 0, zipf: float = 0.8,
                      n_user_groups: int = 0) -> SyntheticInteractions:
    rng = np.random.default_rng(seed)
    p = 1. / np.arange(1, n_items + 1) ** zipf
    cdf = np.cumsum(p / p.sum())
    cdf[-1] = 1.0
    n_draw = int(n_interactions * 1.15)
    u = rng.integers(0, n_users, n_draw, dtype=np.int64)
    i = np.searchsorted(cdf, rng.random(n_draw), side='right').astype(np.int64)
    np.minimum(i, n_items - 1, out=i)
    key = u * n_items + i
    _, first = np.unique(key, return_index=True)
    first.sort()  # keep draw order like DataFrame.drop_duplicates
    first = first[:n_interactions]
    u, i = u[first], i[first]
    r = rng.random(len(u))
    tr, va = r < 0.8, (r >= 0.8) & (r < 0.9)
    te = r >= 0.9
    grp = rng.integers(0, n_user_groups, n_users, dtype=np.int64) if n_user_groups > 0 else None
    return SyntheticInteractions(n_users, n_items, _csr(u[tr], i[tr], n_users, n_items),
                                 _csr(u[va], i[va], n_users, n_items), _csr(u[te], i[te], n_users, n_items), grp,
                                 n_user_groups)


def make_named(name: str, seed: int = 0, n_user_groups: Optional[int] = None) -> SyntheticInteractions:
    U, I, N = SHAPES[name]
    if n_user_groups is None:
        n_user_groups = 2 if name in ('ml1m', 'tiny') else 0  # real ML-1M has 2 gender groups
    return make_interactions(U, I, N, seed=seed, n_user_groups=n_user_groups)


def write_csv_dataset(data: SyntheticInteractions, path: str) -> str:
    """user_idxs.csv, item_idxs.csv, listening_history_{train,val,test}.csv (data/dataset.py:10-23)."""
    import pandas as pd
    os.makedirs(path, exist_ok=True)
    users = pd.DataFrame({'user_idx': np.arange(data.n_users)})
    if data.user_group is not None:
        users['group_idx'] = data.user_group
    users.to_csv(os.path.join(path, 'user_idxs.csv'), index=False)
    pd.DataFrame({'item_idx': np.arange(data.n_items)}).to_csv(os.path.join(path, 'item_idxs.csv'), index=False)
    for split in ('train', 'val', 'test'):
        coo = getattr(data, split).tocoo()
        pd.DataFrame({'user_idx': coo.row.astype(np.int64), 'item_idx': coo.col.astype(np.int64)}).to_csv(
            os.path.join(path, f'listening_history_{split}.csv'), index=False)
    return path
