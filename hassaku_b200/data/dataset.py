"""Datasets behind the reference API (data/dataset.py:26-201): same CSV layout (:10-23), same attributes
(`n_users, n_items, n_user_groups, user_to_user_group, iteration_matrix, sampling_matrix, pop_distribution,
exclude_data`) and the same `__getitem__` contracts, plus `from_interactions` constructors for synthetic data that
never touches disk.  The GPU paths read the CSR matrices directly (uploaded once); `__getitem__` exists for
drop-in compatibility with host-side consumers."""
import logging
import os

import numpy as np
import pandas as pd
import torch
from scipy import sparse as sp
from torch.utils import data


def _csr_from_pairs(u, i, shape, dtype):
    m = sp.csr_matrix((np.ones(len(u), dtype=dtype), (u, i)), shape=shape)
    m.sum_duplicates()
    m.sort_indices()
    return m


class RecDataset(data.Dataset):
    """data/dataset.py:26-86: user_idxs.csv, item_idxs.csv, listening_history_{split}.csv under `data_path`."""

    def __init__(self, data_path: str, split_set: str):
        assert split_set in ['train', 'val', 'test'], f'<{split_set}> is not a valid value for split set!'
        self.data_path, self.split_set = data_path, split_set
        self.n_users = self.n_items = None
        self.user_to_user_group = None  # optional
        self.n_user_groups = 0  # optional
        self.lhs = None
        if data_path is not None:
            self._load_data()
        self.name = 'RecDataset'

    def _load_data(self):
        users = pd.read_csv(os.path.join(self.data_path, 'user_idxs.csv'))
        items = pd.read_csv(os.path.join(self.data_path, 'item_idxs.csv'))
        self.n_users, self.n_items = len(users), len(items)
        if 'group_idx' in users.columns:  # dataset.py:68-72
            grp = users[['user_idx', 'group_idx']].set_index('user_idx').sort_index().group_idx
            self.user_to_user_group = torch.Tensor(grp.to_numpy().copy())
            self.n_user_groups = users.group_idx.nunique()
        self.lhs = self._load_lhs(self.split_set)
        logging.info(f'Loaded {self.split_set}: {self.n_users} users, {self.n_items} items, {len(self.lhs)} interactions')

    def _load_lhs(self, split_set: str):
        return pd.read_csv(os.path.join(self.data_path, f'listening_history_{split_set}.csv'))

    def _set_groups(self, user_group, n_user_groups):
        if user_group is not None:
            self.user_to_user_group = torch.Tensor(np.asarray(user_group))
            self.n_user_groups = int(n_user_groups)

    def __len__(self):
        raise NotImplementedError('RecDataset does not support __len__ or __getitem__. Please use TrainRecDataset for'
                                  'training or FullEvalDataset for evaluation.')

    def __getitem__(self, index):
        raise NotImplementedError('RecDataset does not support __len__ or __getitem__. Please use TrainRecDataset for'
                                  'training or FullEvalDataset for evaluation.')


class TrainRecDataset(RecDataset):
    """data/dataset.py:89-140: COO for iteration, CSR for negative sampling, item popularity distribution."""

    def __init__(self, data_path: str, delete_lhs: bool = True):
        super().__init__(data_path, 'train')
        self.delete_lhs = delete_lhs
        self.iteration_matrix = self.sampling_matrix = self.pop_distribution = None
        if data_path is not None:
            self._prepare_data(self.lhs.user_idx.to_numpy(), self.lhs.item_idx.to_numpy())
            if delete_lhs:
                del self.lhs
        self.name = 'TrainRecDataset'

    @classmethod
    def from_interactions(cls, train_csr, user_group=None, n_user_groups=0):
        self = cls(None)
        self.n_users, self.n_items = train_csr.shape
        coo = train_csr.tocoo()
        self._prepare_data(coo.row, coo.col)
        self._set_groups(user_group, n_user_groups)
        return self

    def _prepare_data(self, u, i):
        self.iteration_matrix = sp.coo_matrix((np.ones(len(u), dtype=np.int16), (u, i)),
                                              shape=(self.n_users, self.n_items))
        self.sampling_matrix = sp.csr_matrix(self.iteration_matrix)
        self.sampling_matrix.sort_indices()
        item_popularity = np.array(self.iteration_matrix.sum(axis=0)).flatten()
        self.pop_distribution = item_popularity / item_popularity.sum()

    def __len__(self):
        return self.iteration_matrix.nnz

    def __getitem__(self, index):
        return (self.iteration_matrix.row[index].astype('int64'), self.iteration_matrix.col[index].astype('int64'), 1.)


class FullEvalDataset(RecDataset):
    """data/dataset.py:143-201: CSR labels of the split + CSR `exclude_data` (train for val; train + val for test)."""

    def __init__(self, data_path: str, split_set: str, delete_lhs: bool = True):
        super().__init__(data_path, split_set)
        self.delete_lhs = delete_lhs
        self.idx_to_user = None
        self.iteration_matrix = self.exclude_data = None
        if data_path is not None:
            shape = (self.n_users, self.n_items)
            self.iteration_matrix = _csr_from_pairs(self.lhs.user_idx, self.lhs.item_idx, shape, np.int16)
            tr = self._load_lhs('train')
            self.exclude_data = _csr_from_pairs(tr.user_idx, tr.item_idx, shape, bool)
            if split_set == 'test':
                va = self._load_lhs('val')
                self.exclude_data = sp.csr_matrix(self.exclude_data + _csr_from_pairs(va.user_idx, va.item_idx, shape, bool))
                self.exclude_data.sort_indices()
            if delete_lhs:
                del self.lhs
        self.name = 'FullEvalDataset'

    @classmethod
    def from_interactions(cls, labels_csr, exclude_csr, split_set='val', user_group=None, n_user_groups=0):
        self = cls(None, split_set)
        self.n_users, self.n_items = labels_csr.shape
        self.iteration_matrix = sp.csr_matrix(labels_csr)
        self.iteration_matrix.sort_indices()
        self.exclude_data = sp.csr_matrix(exclude_csr)
        self.exclude_data.sort_indices()
        self._set_groups(user_group, n_user_groups)
        return self

    def __len__(self):
        return self.n_users

    def __getitem__(self, user_index):
        return user_index, np.arange(self.n_items), self.iteration_matrix[user_index].toarray().squeeze().astype('float32')
