"""GPU: the device negative sampler (bit-exact vs the numpy contract in oracle/philox.py) and the Trainer end to end
(reference control flow: validation before training, per-epoch validation, best-model checkpoint, patience)."""
import os
import tempfile

import numpy as np
import pytest
import torch
from scipy import sparse as sp

pytestmark = pytest.mark.gpu


def _tiny():
    from hassaku_b200.data.synthetic import make_interactions
    return make_interactions(300, 200, 6000, seed=0, n_user_groups=2)


@pytest.mark.parametrize('N,distinct', [(20, True), (50, True), (7, False), (100, True)])
def test_sampler_bit_exact_vs_oracle_contract(N, distinct):
    from oracle import philox as P
    from hassaku_b200 import _C
    data = _tiny()
    csr = data.train
    rng = np.random.RandomState(0)
    B = 96
    u = rng.randint(0, 300, B).astype(np.int64)
    pos = np.array([csr[x].indices[0] if csr[x].nnz else 0 for x in u], dtype=np.int64)
    indptr = torch.from_numpy(csr.indptr.astype(np.int64)).cuda()
    indices = torch.from_numpy(csr.indices.astype(np.int32)).cuda()
    for seed, step in [(64, 0), (64, 7), ((1 << 40) + 3, (1 << 33) + 5)]:
        out = torch.empty((B, N + 1), dtype=torch.int64, device='cuda')
        _C.sample_negatives(torch.from_numpy(u).cuda(), torch.from_numpy(pos).cuda(), N, 200, 300, indptr, indices, seed,
                            step, out, distinct)
        ref = P.sample_negatives(u, N, 200, csr.indptr, csr.indices, seed, step, distinct)
        got = out.cpu().numpy()
        np.testing.assert_array_equal(got[:, 0], pos)
        np.testing.assert_array_equal(got[:, 1:], ref)


def test_sampler_semantics_at_scale():
    """ML-1M shape, B = 8192, N = 50: no train item among the negatives, distinct rows, uniform marginal."""
    from hassaku_b200.data.dataset import TrainRecDataset
    from hassaku_b200.data.dataloader import NegativeSampler, TrainDataLoader
    from hassaku_b200.data.synthetic import make_named
    data = make_named('ml1m')
    ds = TrainRecDataset.from_interactions(data.train)
    dl = TrainDataLoader(NegativeSampler(ds, n_neg=50), ds, batch_size=8192, shuffle=True)
    assert len(dl) == -(-data.train.nnz // 8192)
    it = iter(dl)
    counts = np.zeros(data.n_items)
    for _ in range(4):
        u, i, lab = next(it)
        assert u.dtype == torch.int64 and i.shape == (8192, 51) and lab.dtype == torch.float64
        assert float(lab[:, 0].min()) == 1.0 and float(lab[:, 1:].abs().max()) == 0.0
        u_np, i_np = u.cpu().numpy(), i.cpu().numpy()
        keys = set((data.train.tocoo().row.astype(np.int64) * data.n_items + data.train.tocoo().col).tolist()) \
            if _ == 0 else keys
        assert all((int(a) * data.n_items + int(b)) in keys for a, b in zip(u_np[:512], i_np[:512, 0]))  # positives
        neg_keys = (u_np[:, None] * data.n_items + i_np[:, 1:]).ravel()
        assert not np.isin(neg_keys, np.fromiter(keys, dtype=np.int64)).any()
        assert all(len(set(r)) == 50 for r in i_np[:256, 1:])
        counts += np.bincount(i_np[:, 1:].ravel(), minlength=data.n_items)
    # popular (low-id) items are excluded more often (they are train items of many users); the rest is flat
    tail = counts[2000:]
    assert abs(tail.std() / tail.mean() - 1 / np.sqrt(tail.mean())) < 0.05
    assert int(dl.status.item()) == 0


class _ListLoader:
    """A 'loader' replaying captured batches (what the parity tests feed both implementations)."""

    def __init__(self, batches, dataset=None):
        self.batches, self.dataset = batches, dataset

    def __iter__(self):
        return iter(self.batches)

    def __len__(self):
        return len(self.batches)


class _EvalLoader:
    def __init__(self, dataset, batch_size):
        self.dataset, self.batch_size = dataset, batch_size


def _conf(path, **kw):
    c = {'device': 'cuda', 'lr': 5e-3, 'wd': 1e-5, 'optimizer': 'adamw', 'n_epochs': 3, 'optimizing_metric': 'ndcg@10',
         'max_patience': 2, 'model_path': path, 'running_settings': {'use_wandb': False, 'batch_verbose': False}}
    c.update(kw)
    return c


def test_trainer_fit_matches_oracle_on_captured_batches():
    """Same captured batches through hassaku_b200.Trainer and through the oracle's trainer.py:128-148 restatement:
    epoch losses and validation metrics after every epoch agree (free-running, 3 epochs x 12 steps)."""
    from oracle import mf_oracle as O
    from hassaku_b200.algorithms.sgd_alg import SGDMatrixFactorization
    from hassaku_b200.data.dataset import FullEvalDataset
    from hassaku_b200.train.rec_losses import RecBayesianPersonalizedRankingLoss
    from hassaku_b200.train.trainer import Trainer
    data = _tiny()
    U, I, d, B, N = 300, 200, 18, 128, 10
    rng = np.random.RandomState(3)
    coo = data.train.tocoo()
    batches = []
    for _ in range(12):
        sel = rng.randint(0, coo.nnz, B)
        i = np.column_stack([coo.col[sel], rng.randint(0, I, (B, N))]).astype(np.int64)
        batches.append((torch.from_numpy(coo.row[sel].astype(np.int64)), torch.from_numpy(i), O.make_labels(B, N + 1)))
    torch.manual_seed(64)
    ref = O.OracleMF(U, I, d, use_item_bias=True)
    torch.manual_seed(64)
    model = SGDMatrixFactorization(U, I, d, use_item_bias=True)
    with torch.no_grad():  # O(1) scores so that three epochs actually move the metrics
        for m_ in (ref, model):
            m_.user_embeddings.weight.mul_(40.)
            m_.item_embeddings.weight.mul_(40.)
    ds = FullEvalDataset.from_interactions(data.val, data.train, 'val', data.user_group, 2)
    with tempfile.TemporaryDirectory() as tmp:
        tr = Trainer(model, _ListLoader(batches), _EvalLoader(ds, 128), RecBayesianPersonalizedRankingLoss(), _conf(tmp))
        u2g = torch.from_numpy(data.user_group).float()
        otr = O.OracleTrainer(ref, 'bpr', 5e-3, 1e-5, 'adamw', neg_train=N)
        ref_metrics = [O.evaluate(ref, data.val, data.train, 128, 2, u2g)]
        ref_losses = []
        for ep in range(3):
            tot = 0.
            for u, i, lab in batches:
                tot += float(otr.step(u, i, lab)['loss'])
            ref_losses.append(tot / len(batches))
            ref_metrics.append(O.evaluate(ref, data.val, data.train, 128, 2, u2g))
        logs = []
        tr._report = lambda log_dict, epoch: logs.append((epoch, dict(log_dict)))
        best = tr.fit()
        assert os.path.exists(os.path.join(tmp, 'model.pth'))
    assert [e for e, _ in logs] == [-1, 0, 1, 2]
    for (ep, log), rm in zip(logs, ref_metrics):
        for k, v in rm.items():
            assert abs(log[k] - v) < 2e-3, (ep, k, log[k], v)   # free-running: a tie flip moves a metric by 1/300/k
        if ep >= 0:
            assert abs(log['epoch_train_loss'] - ref_losses[ep]) <= 2e-5 * abs(ref_losses[ep])
            assert log['epoch_train_rec_loss'] == log['epoch_train_loss'] and log['epoch_train_reg_loss'] == 0.0
    best_ref = max(range(4), key=lambda e: (ref_metrics[e]['ndcg@10'], -e))
    assert best['best_epoch'] == best_ref - 1
    assert best['max_optimizing_metric'] == tr.best_value == best['ndcg@10']
    assert len([k for k in best if k.startswith('group_')]) == 24


def test_trainer_with_device_loader_learns_and_stops_on_patience():
    from hassaku_b200.algorithms.sgd_alg import SGDMatrixFactorization
    from hassaku_b200.data.dataloader import NegativeSampler, TrainDataLoader
    from hassaku_b200.data.dataset import FullEvalDataset, TrainRecDataset
    from hassaku_b200.train.rec_losses import RecSampledSoftmaxLoss
    from hassaku_b200.train.trainer import Trainer
    data = _tiny()
    tds = TrainRecDataset.from_interactions(data.train, data.user_group, 2)
    dl = TrainDataLoader(NegativeSampler(tds, n_neg=20), tds, batch_size=256, shuffle=True)
    ds = FullEvalDataset.from_interactions(data.val, data.train, 'val', data.user_group, 2)
    torch.manual_seed(1)
    model = SGDMatrixFactorization(300, 200, 32, use_item_bias=True)
    loss = RecSampledSoftmaxLoss(n_items=200, train_neg_strategy='uniform', neg_train=20)
    with tempfile.TemporaryDirectory() as tmp:
        tr = Trainer(model, dl, _EvalLoader(ds, 300), loss, _conf(tmp, n_epochs=12, max_patience=3, lr=2e-2))
        first = tr.val()
        best = tr.fit()
        assert best['ndcg@10'] > first['ndcg@10'] + 0.01           # it learns
        assert -1 <= best['best_epoch'] < 12
        # the checkpoint on disk is the best model, in the reference's state_dict format
        sd = torch.load(os.path.join(tmp, 'model.pth'), map_location='cpu')
        assert set(sd) == {'user_embeddings.weight', 'item_embeddings.weight', 'item_bias.weight'}
        m2 = SGDMatrixFactorization(300, 200, 32, use_item_bias=True)
        m2.load_model_from_path(tmp)
        tr2 = Trainer(m2, dl, _EvalLoader(ds, 300), loss, _conf(tmp))
        assert abs(tr2.val()['ndcg@10'] - best['ndcg@10']) < 1e-7


def test_popular_sampler_bit_exact_vs_oracle_and_distribution():
    from oracle import philox as P
    from hassaku_b200 import _C
    from hassaku_b200.data.dataset import TrainRecDataset
    from hassaku_b200.data.dataloader import NegativeSampler, TrainDataLoader
    data = _tiny()
    ds = TrainRecDataset.from_interactions(data.train)
    sampler = NegativeSampler(ds, n_neg=12, neg_sampling_strategy='popular', squashing_factor_pop_sampling=0.75)
    cdf = sampler.popularity_cdf()
    np.testing.assert_array_equal(cdf, P.popularity_cdf_u64(ds.pop_distribution, 0.75))
    assert cdf[-1] == np.uint64(2 ** 64 - 1) and (np.diff(cdf.astype(np.float64)) >= 0).all()
    dl = TrainDataLoader(sampler, ds, batch_size=64, shuffle=False, seed=5)
    u, i, _ = next(iter(dl))
    ref = P.sample_negatives(u.cpu().numpy(), 12, 200, data.train.indptr, data.train.indices, 5, 0, True, pop_cdf=cdf)
    np.testing.assert_array_equal(i.cpu().numpy()[:, 1:], ref)
    # marginal follows pop^alpha over the allowed items (coarsely: popular items are drawn more often)
    sampler2 = NegativeSampler(ds, n_neg=50, neg_sampling_strategy='popular', distinct_in_row=False)
    dl2 = TrainDataLoader(sampler2, ds, batch_size=2048, shuffle=True, seed=1)
    cnt = np.zeros(200)
    for b, (u, i, _) in enumerate(dl2):
        cnt += np.bincount(i[:, 1:].flatten().cpu().numpy(), minlength=200)
        if b == 1:
            break
    pop = ds.pop_distribution
    top, bottom = np.argsort(-pop)[:20], np.argsort(-pop)[-60:]
    assert cnt[top].mean() > 2 * cnt[bottom].mean()


def test_trainer_adagrad_option_runs_on_the_fused_path():
    from hassaku_b200.algorithms.sgd_alg import SGDMatrixFactorization
    from hassaku_b200.data.dataloader import NegativeSampler, TrainDataLoader
    from hassaku_b200.data.dataset import FullEvalDataset, TrainRecDataset
    from hassaku_b200.train.optim import DenseAdagrad
    from hassaku_b200.train.rec_losses import RecBinaryCrossEntropy
    from hassaku_b200.train.trainer import Trainer
    data = _tiny()
    tds = TrainRecDataset.from_interactions(data.train, data.user_group, 2)
    dl = TrainDataLoader(NegativeSampler(tds, n_neg=8), tds, batch_size=256, shuffle=True)
    ds = FullEvalDataset.from_interactions(data.val, data.train, 'val', data.user_group, 2)
    torch.manual_seed(2)
    model = SGDMatrixFactorization(300, 200, 16, use_item_bias=True)
    with tempfile.TemporaryDirectory() as tmp:
        tr = Trainer(model, dl, _EvalLoader(ds, 300), RecBinaryCrossEntropy(), _conf(tmp, optimizer='adagrad', lr=5e-2, n_epochs=4))
        assert isinstance(tr.optimizer, DenseAdagrad)
        first = tr.val()['ndcg@10']
        best = tr.fit()
        assert best['ndcg@10'] >= first
