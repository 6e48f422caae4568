"""User / item distributions over tags and popularity buckets for the calibration metrics
(reference data/data_utils.py:378-499, consumed by FullEvaluatorCalibrationDecorator in sweep_test.py:65-69).

Same CSV layout (`<dataset>/processed_dataset/{item_idxs,user_idxs,tag_idxs,item_tag_idxs,listening_history_train}.csv`),
same fp32 arithmetic and the same `(user_matrix, item_matrix)` return order; the per-item Python loop of the popularity
bucketing is a cumulative sum here.  `*_from_csr` variants take the train matrix directly (synthetic data)."""
import os
from typing import Tuple

import numpy as np
import scipy.sparse as sp
import torch


def _train_csr(path_to_dataset_folder: str, dtype) -> Tuple[sp.csr_matrix, int, int]:
    import pandas as pd
    base = os.path.join(path_to_dataset_folder, 'processed_dataset')
    n_items = len(pd.read_csv(os.path.join(base, 'item_idxs.csv')))
    n_users = len(pd.read_csv(os.path.join(base, 'user_idxs.csv')))
    tr = pd.read_csv(os.path.join(base, 'listening_history_train.csv'))[['user_idx', 'item_idx']]
    m = sp.csr_matrix((np.ones(len(tr), dtype=dtype), (tr.user_idx.to_numpy(), tr.item_idx.to_numpy())),
                      shape=(n_users, n_items))
    return m, n_users, n_items


def tag_matrices_from_csr(train: sp.csr_matrix, item_idx: np.ndarray, tag_idx: np.ndarray, n_tags: int,
                          alpha_smoothening: float = .01) -> Tuple[torch.Tensor, torch.Tensor]:
    """item x tag: 1 / (#tags of the item) on the item's tags, zero rows for untagged items (data_utils.py:407-413);
    user x tag: mean of the item rows over the user's train items, smoothed with alpha / n_tags (eq. 7, :421-427)."""
    assert 0 <= alpha_smoothening <= 1, 'Alpha value out of bounds'
    n_items = train.shape[1]
    tag = np.zeros((n_items, n_tags), dtype=np.float32)
    tag[np.asarray(item_idx), np.asarray(tag_idx)] = 1.
    tot = tag.sum(-1, dtype=np.float32)
    with np.errstate(invalid='ignore', divide='ignore'):
        tag = np.where(tot[:, None] > 0, tag / tot[:, None], np.float32(0.)).astype(np.float32)
        freq = np.asarray(train.astype(np.int16) @ tag, dtype=np.float32)
        freq /= np.asarray(train.sum(-1))                       # users without train items: 0 / 0 = nan, like the reference
    freq = (alpha_smoothening / n_tags + (1 - alpha_smoothening) * freq).astype(np.float32)
    return torch.from_numpy(freq), torch.from_numpy(tag)


def pop_matrices_from_csr(train: sp.csr_matrix, alpha_smoothening: float = .01) -> Tuple[torch.Tensor, torch.Tensor]:
    """item x {head, middle, tail}: items sorted by decreasing train popularity; an item is 'head' while the cumulative
    popularity mass (including it) is < 0.2, 'middle' while < 0.8, else 'tail' (data_utils.py:458-487);
    user x bucket: share of the user's train items per bucket, smoothed with alpha / 3 (:491-497)."""
    assert 0 <= alpha_smoothening <= 1, 'Alpha value out of bounds'
    train = train.astype(np.float32)
    n_items = train.shape[1]
    pop = np.asarray(train.sum(0), dtype=np.float32).ravel()
    pop /= pop.sum(dtype=np.float32)
    order = np.argsort(-pop)                                     # same call as the reference: same tie order
    cum = np.cumsum(pop[order], dtype=np.float32)                # sequential fp32 accumulation, like its `+=` loop
    bucket = np.where(cum < np.float32(0.2), 0, np.where(cum < np.float32(0.8), 1, 2))
    item_pop = np.zeros((n_items, 3), dtype=np.float32)
    item_pop[order, bucket] = 1.
    with np.errstate(invalid='ignore', divide='ignore'):
        user_pop = np.asarray(train @ item_pop, dtype=np.float32)
        user_pop /= user_pop.sum(-1, dtype=np.float32)[:, None]
    user_pop = (alpha_smoothening / 3 + (1 - alpha_smoothening) * user_pop).astype(np.float32)
    return torch.from_numpy(user_pop), torch.from_numpy(item_pop)


def build_user_and_item_tag_matrix(path_to_dataset_folder: str, alpha_smoothening: float = .01):
    """data_utils.py:378-429 — returns (user_tag_matrix [U, T], item_tag_matrix [I, T])."""
    import pandas as pd
    base = os.path.join(path_to_dataset_folder, 'processed_dataset')
    n_tags = len(pd.read_csv(os.path.join(base, 'tag_idxs.csv')))
    it = pd.read_csv(os.path.join(base, 'item_tag_idxs.csv'))
    train, _, _ = _train_csr(path_to_dataset_folder, np.int16)
    return tag_matrices_from_csr(train, it.item_idx.to_numpy(), it.tag_idx.to_numpy(), n_tags, alpha_smoothening)


def build_user_and_item_pop_matrix(path_to_dataset_folder: str, alpha_smoothening: float = .01):
    """data_utils.py:432-499 — returns (user_pop_matrix [U, 3], item_pop_matrix [I, 3])."""
    train, _, _ = _train_csr(path_to_dataset_folder, np.float32)
    return pop_matrices_from_csr(train, alpha_smoothening)
