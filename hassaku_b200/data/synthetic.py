"""Seeded synthetic implicit-feedback data of the shapes BASELINE.json names.

Users uniform, items Zipf-like p(i) ∝ (rank+1)^-0.8, de-duplicated (u, i) pairs, per-interaction
uniform r in [0,1) -> 80/10/10 train/val/test (mirrors the ratio split of the reference's
data/data_utils.py:241-277).  `write_csv_dataset` emits the 5-file layout the reference's loaders
read (data/dataset.py:10-23) so the *unmodified* reference can consume the same data.
"""
from __future__ import annotations

import math
import os
from dataclasses import dataclass
from typing import Optional

import numpy as np
from scipy import sparse as sp

SHAPES = {
    # name: (n_users, n_items, n_interactions)
    'ml1m': (6040, 3706, 1_000_000),
    'ml10m': (69_878, 10_677, 10_000_000),
    'lfm2b': (2_000_000, 1_000_000, 200_000_000),          # cfg4
    'eval10m': (10_000_000, 1_000_000, 1_000_000_000),     # cfg5: 10 M users, ~80 train + ~10 val ids per user
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


# ---- the large shapes (cfg4 / cfg5) are generated ON THE DEVICE, per rank: the reference's loaders cannot hold them ----
class DeviceInteractions:
    """The interactions of the users ONE rank owns (u % world == rank) of a synthetic data set, as device CSR over the
    GLOBAL user id space (rows of other ranks' users are empty): `train` / `val` = (indptr int64 [n_users + 1], indices
    int32 sorted per row), and the train COO list (`rows`, `cols`, int64) that an epoch iterates."""

    def __init__(self, n_users, n_items, world, rank, train, val, rows, cols):
        self.n_users, self.n_items, self.world, self.rank = n_users, n_items, world, rank
        self.train, self.val, self.rows, self.cols = train, val, rows, cols


def make_device_interactions(n_users: int, n_items: int, n_interactions: int, device, world: int = 1, rank: int = 0,
                             seed: int = 0, zipf: float = 0.8, keep_coo: bool = True, chunk_draws: int = 48_000_000):
    """Same distributions as make_interactions (users uniform, items Zipf 0.8, de-duplicated pairs, 80 / 10 / 10 split by a
    per-pair uniform draw) generated with torch on `device` for the users of `rank`, in chunks of users so that the
    sort that de-duplicates never holds more than `chunk_draws` keys.  Deterministic in (seed, world, rank)."""
    import torch
    dev = torch.device(device)
    local_users = torch.arange(rank, n_users, world, device=dev, dtype=torch.int64)
    nl = int(local_users.numel())
    n_local = int(round(n_interactions * nl / max(n_users, 1)))
    p = 1.0 / torch.arange(1, n_items + 1, device=dev, dtype=torch.float64) ** zipf
    cdf = torch.cumsum(p / p.sum(), 0)
    cdf[-1] = 1.0
    n_chunks = max(1, math.ceil(n_local / chunk_draws))
    per_chunk_users = math.ceil(nl / n_chunks)
    parts = {'train': [], 'val': []}
    coo_u, coo_i = [], []
    cnt = {k: torch.zeros(n_users, dtype=torch.int64, device=dev) for k in parts}
    for c in range(n_chunks):
        c0, c1 = c * per_chunk_users, min(nl, (c + 1) * per_chunk_users)
        if c1 <= c0:
            break
        gen = torch.Generator(device=dev)
        gen.manual_seed((seed * 1_000_003 + rank) * 4099 + c)
        n_draw = int(round(n_local * (c1 - c0) / nl))
        u = local_users[c0 + torch.randint(0, c1 - c0, (n_draw,), device=dev, generator=gen)]
        i = torch.searchsorted(cdf, torch.rand(n_draw, device=dev, dtype=torch.float64, generator=gen), right=True)
        i.clamp_(max=n_items - 1)
        key = torch.unique(u * n_items + i, sorted=True)         # sorted by (user, item), duplicates dropped
        del u, i
        r = torch.rand(key.numel(), device=dev, generator=gen)
        for name, sel in (('train', r < 0.8), ('val', (r >= 0.8) & (r < 0.9))):
            k = key[sel]
            uu = torch.div(k, n_items, rounding_mode='floor')
            parts[name].append((k - uu * n_items).to(torch.int32))
            cnt[name] += torch.bincount(uu, minlength=n_users)
            if name == 'train' and keep_coo:
                coo_u.append(uu)
                coo_i.append(k - uu * n_items)
        del key, r
    out = {}
    for name in parts:
        indptr = torch.zeros(n_users + 1, dtype=torch.int64, device=dev)
        torch.cumsum(cnt[name], 0, out=indptr[1:])
        # chunks cover ascending, disjoint ranges of this rank's users and every chunk is sorted by (user, item): the
        # concatenation is the CSR index array
        idx = torch.cat(parts[name]) if parts[name] else torch.zeros(0, dtype=torch.int32, device=dev)
        out[name] = (indptr, idx)
    rows = torch.cat(coo_u) if coo_u else None
    cols = torch.cat(coo_i) if coo_i else None
    return DeviceInteractions(n_users, n_items, world, rank, out['train'], out['val'], rows, cols)
