"""Training loader behind the reference names (data/dataloader.py:17-129): `NegativeSampler` + `TrainDataLoader`
yielding `(u_idxs int64 [B], i_idxs int64 [B, 1+N], labels float64 [B, 1+N])` — but device resident: the train COO
list lives in HBM, an epoch is a device permutation of it, and the negatives come from hsk_sample_negatives (Philox,
rejection against the user's train CSR row) instead of a per-row numpy loop in worker processes (~8 k samples/s in
the reference).  Nothing crosses PCIe per batch."""
import logging
import math

import numpy as np
import torch

from hassaku_b200 import _C
from hassaku_b200.data.dataset import TrainRecDataset


class InteractionSampler:
    pass


class NegativeSampler(InteractionSampler):
    """data/dataloader.py:17-64: parameters of the negative sampling: 'uniform' or 'popular' (item ~ pop^alpha)."""

    def __init__(self, train_dataset: TrainRecDataset, n_neg: int = 10, neg_sampling_strategy: str = 'uniform',
                 squashing_factor_pop_sampling: float = 1., distinct_in_row: bool = True):
        assert n_neg > 0, 'Number of negatives should be > 0!'
        assert neg_sampling_strategy in ['uniform', 'popular'], \
            f'<{neg_sampling_strategy}> is not a valid negative sampling strategy!'
        assert squashing_factor_pop_sampling >= 0, 'Squashing factor for popularity sampling should be positive!'
        self.dataset = train_dataset
        self.n_neg = n_neg
        self.neg_sampling_strategy = neg_sampling_strategy
        self.squashing_factor_pop_sampling = squashing_factor_pop_sampling
        self.distinct_in_row = distinct_in_row
        self.n_items = train_dataset.n_items
        self.pop_distribution = train_dataset.pop_distribution.copy()
        self.name = 'NegativeSampler'
        logging.info(f'Built {self.name} module: n_neg={n_neg}, strategy={neg_sampling_strategy}')

    def popularity_cdf(self) -> np.ndarray:
        """64-bit fixed-point CDF of pop^alpha (what hsk_sample_negatives consumes for the 'popular' strategy)."""
        p = np.power(np.asarray(self.pop_distribution, dtype=np.float64), self.squashing_factor_pop_sampling)
        p = p / p.sum()
        c = np.cumsum(p)
        c = c / c[-1]
        out = np.array([min(int(x * 18446744073709551616.0), 18446744073709551615) for x in c], dtype=np.uint64)
        out[-1] = np.uint64(18446744073709551615)
        return out


class TrainDataLoader:
    """Iterable with the reference TrainDataLoader's batch contract.  `device` defaults to the current CUDA device;
    `seed` + the running batch counter key the Philox stream, so a (seed, epoch, batch) triple reproduces a batch
    exactly regardless of how the loader is driven (the reference's negatives depend on the worker count)."""

    def __init__(self, interaction_sampler: InteractionSampler, dataset: TrainRecDataset, batch_size: int = 1,
                 shuffle: bool = False, drop_last: bool = False, device=None, seed: int = 64, **_ignored):
        if not isinstance(interaction_sampler, NegativeSampler):
            raise ValueError('Invalid Interaction Sampler')
        if not torch.cuda.is_available():
            raise _C.HskError('hassaku_b200.TrainDataLoader needs a CUDA device (no CPU path)')
        self.interaction_sampler, self.dataset = interaction_sampler, dataset
        self.batch_size, self.shuffle, self.drop_last, self.seed = int(batch_size), shuffle, drop_last, int(seed)
        self.device = torch.device(device if device is not None else 'cuda')
        coo = dataset.iteration_matrix
        self.rows = torch.from_numpy(np.asarray(coo.row, dtype=np.int64)).to(self.device)
        self.cols = torch.from_numpy(np.asarray(coo.col, dtype=np.int64)).to(self.device)
        csr = dataset.sampling_matrix
        if not csr.has_sorted_indices:
            csr = csr.sorted_indices()
        self.indptr = torch.from_numpy(csr.indptr.astype(np.int64)).to(self.device)
        self.indices = torch.from_numpy(csr.indices.astype(np.int32)).to(self.device)
        self.status = torch.zeros(1, dtype=torch.int32, device=self.device)
        self.step = 0
        self._labels = {}
        self.pop_cdf = None
        if interaction_sampler.neg_sampling_strategy == 'popular':
            self.pop_cdf = torch.from_numpy(interaction_sampler.popularity_cdf().view(np.int64)).to(self.device)
        self._gen = torch.Generator(device=self.device)
        self._gen.manual_seed(self.seed)

    def __len__(self):
        n = len(self.dataset)
        return n // self.batch_size if self.drop_last else math.ceil(n / self.batch_size)

    def labels_for(self, B: int) -> torch.Tensor:
        """float64 [B, 1+N], column 0 = 1 (data/dataloader.py:126-128); constant, so built once per shape."""
        if B not in self._labels:
            lab = torch.zeros((B, self.interaction_sampler.n_neg + 1), dtype=torch.float64, device=self.device)
            lab[:, 0] = 1.
            self._labels[B] = lab
        return self._labels[B]

    def sample_batch(self, u_idxs: torch.Tensor, pos_idxs: torch.Tensor, step: int) -> torch.Tensor:
        s = self.interaction_sampler
        i_idxs = torch.empty((len(u_idxs), s.n_neg + 1), dtype=torch.int64, device=self.device)
        _C.sample_negatives(u_idxs, pos_idxs, s.n_neg, self.dataset.n_items, self.dataset.n_users, self.indptr,
                            self.indices, self.seed, step, i_idxs, s.distinct_in_row, self.status, pop_cdf=self.pop_cdf)
        return i_idxs

    def check_status(self):
        """One host sync: raise if a sampler launch since the last check met a bad user index or hit its round cap."""
        st = int(self.status.item())
        if st:
            self.status.zero_()
            what = []
            if st & _C.STATUS_BAD_INDEX:
                what.append('a user index outside [0, n_users)')
            if st & _C.STATUS_SAMPLER_ROUNDS:
                what.append('a row that still held train positives / duplicates after the round cap (neg_train larger '
                            'than the number of admissible items of a user?)')
            raise _C.HskError('hsk_sample_negatives reported ' + ' and '.join(what))

    def __iter__(self):
        n = len(self.dataset)
        order = torch.randperm(n, device=self.device, generator=self._gen) if self.shuffle else \
            torch.arange(n, device=self.device)
        for b in range(len(self)):
            sel = order[b * self.batch_size:(b + 1) * self.batch_size]
            u_idxs, pos = self.rows[sel], self.cols[sel]
            i_idxs = self.sample_batch(u_idxs, pos, self.step)
            self.step += 1
            yield u_idxs, i_idxs, self.labels_for(len(sel))
        self.check_status()      # once per epoch, like model.check_status()


class EvalLoader:
    """Stand-in for `DataLoader(FullEvalDataset(...), batch_size=...)` (data/data_utils.py:348-368): the fused evaluator
    only needs `.dataset` (its CSR matrices) and `.batch_size`; iterating yields the reference's dense batches
    `(u_idxs, arange(I), y_true)` for host-side consumers."""

    def __init__(self, dataset, batch_size: int = 64):
        self.dataset, self.batch_size = dataset, int(batch_size)

    def __len__(self):
        return math.ceil(len(self.dataset) / self.batch_size)

    def __iter__(self):
        n, I = len(self.dataset), self.dataset.n_items
        for s in range(0, n, self.batch_size):
            u = np.arange(s, min(s + self.batch_size, n))
            y = self.dataset.iteration_matrix[u].toarray().astype('float32')
            yield torch.from_numpy(u.astype(np.int64)), torch.arange(I).repeat(len(u), 1), torch.from_numpy(y)
