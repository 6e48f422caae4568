"""TEST INFRASTRUCTURE ONLY — generates tests/golden/*.npz by running the UNMODIFIED reference
(karapostK/hassaku, imported from /root/reference via oracle/ref_shim.py) on seeded inputs.

Run in the build container only:   python -m oracle.make_golden
The fixtures travel to the GPU box; the reference does not.

Fixtures
  train_{bpr,ssm,bce}.npz  reference SGDMatrixFactorization + rec loss + torch.optim.AdamW/Adam, 3 teacher-forced
                           steps on captured (u_idxs, i_idxs): initial weights, per-step scores, loss, dL/dscores,
                           dense grads, post-step params and optimizer state.
  loader_batches.npz       batches produced by the reference TrainDataLoader (negative sampling semantics).
  eval_tiny.npz            reference evaluate_recommender_algorithm(FullEvaluator) on the 'tiny' synthetic
                           dataset through the reference FullEvalDataset: metric dict, masked scores, top-100 ids.
  metrics_kat.npz          the 15 known answers of framework_tests/eval/test_metrics.py evaluated by the
                           reference's own metric functions.
  calibration_kat.npz      reference FullEvaluatorCalibrationDecorator ('tag' + 'pop', nested) over two dense batches:
                           aggregated dict (2 user groups) and per-user vectors.
  calibration_matrices.npz reference build_user_and_item_{tag,pop}_matrix on the tiny dataset (CSV layout + random tags).
  baseline_init.npz        reference SGDBaseline: seeded initial weights, names / shapes, scores of one batch.
  train_baseline_bce.npz   reference SGDBaseline + bce + AdamW, 3 teacher-forced steps (same layout as train_*.npz).

`python -m oracle.make_golden [calibration] [calibration_matrices] [baseline]` regenerates only the named groups;
without arguments everything (the run is deterministic: regenerated files are identical to the committed ones).
"""
import os
import sys
import tempfile

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
GOLD = os.path.join(ROOT, 'tests', 'golden')

from oracle import ref_shim  # noqa: E402

TINY = dict(n_users=300, n_items=200, n_interactions=6000, seed=0, n_user_groups=2)


def tiny_data():
    from hassaku_b200.data.synthetic import make_interactions
    return make_interactions(**TINY)


def gen_train(name, loss_kind, d, use_user_bias, use_item_bias, use_global_bias, optimizer, B, N, lr, wd, n_steps=3,
              model_kind='mf'):
    from algorithms.sgd_alg import SGDBaseline, SGDMatrixFactorization
    from train.rec_losses import RecommenderSystemLossesEnum
    data = tiny_data()
    U, I = data.n_users, data.n_items
    torch.manual_seed(64)
    rng = np.random.RandomState(64)
    if model_kind == 'baseline':
        model = SGDBaseline(U, I)
    else:
        model = SGDMatrixFactorization(U, I, d, use_user_bias, use_item_bias, use_global_bias)
    # larger-than-init weights so that scores/gradients are O(1) and the comparison is meaningful
    with torch.no_grad():
        for p in model.parameters():
            p.copy_(torch.randn_like(p) * (0.3 if model_kind == 'baseline' else 1.0 / np.sqrt(d) if p.shape[-1] == d else 0.1))
    conf = {'train_neg_strategy': 'uniform', 'neg_train': N}

    class _DS:
        n_items = I

    loss_obj = RecommenderSystemLossesEnum[loss_kind].value.build_from_conf(conf, _DS())
    opt_cls = {'adamw': torch.optim.AdamW, 'adam': torch.optim.Adam, 'adagrad': torch.optim.Adagrad}[optimizer]
    opt = opt_cls(model.parameters(), lr=lr, weight_decay=wd)
    out = {'meta_loss': loss_kind, 'meta_optimizer': optimizer, 'meta_dims': np.array([U, I, d or 0, B, N]),
           'meta_hparams': np.array([lr, wd]),
           'meta_flags': np.array([use_user_bias, use_item_bias, use_global_bias])}
    for n, p in model.state_dict().items():
        out[f'init/{n}'] = p.detach().numpy().copy()
    coo = data.train.tocoo()
    for s in range(n_steps):
        sel = rng.randint(0, coo.nnz, B)
        if s == 1:
            sel[:8] = sel[0]  # force duplicate users/positives in one batch
        u = coo.row[sel].astype(np.int64)
        pos = coo.col[sel].astype(np.int64)
        neg = rng.randint(0, I, (B, N)).astype(np.int64)
        if s == 1:
            neg[:, 1] = neg[:, 0]  # duplicate negatives within rows
        i = np.column_stack([pos, neg])
        labels = np.zeros_like(i, dtype=float)
        labels[:, 0] = 1.
        u_t, i_t, l_t = torch.from_numpy(u), torch.from_numpy(i), torch.from_numpy(labels)
        # --- train/trainer.py:133-148 on the real reference objects ---
        logits = model(u_t, i_t)
        # hook (not retain_grad): sampled-softmax modifies `logits` in place (rec_losses.py:134) and a hook
        # registered before that still receives dL/d(model output), which is what the scatter kernel consumes
        cap = {}
        logits.register_hook(lambda gr: cap.__setitem__('g', gr.detach().clone()))
        scores = logits.detach().clone()
        loss = loss_obj.compute_loss(logits, l_t)
        reg = model.get_and_reset_other_loss()['reg_loss']
        total = loss + reg
        total.backward()
        out[f's{s}/u_idxs'], out[f's{s}/i_idxs'] = u, i
        out[f's{s}/scores'] = scores.numpy()
        out[f's{s}/loss'] = loss.detach().numpy()
        out[f's{s}/dscores'] = cap['g'].numpy().copy()
        for n, p in model.named_parameters():
            out[f's{s}/grad/{n}'] = p.grad.numpy().copy()
        opt.step()
        opt.zero_grad()
        for n, p in model.named_parameters():
            out[f's{s}/param/{n}'] = p.detach().numpy().copy()
            st = opt.state[p]
            if 'exp_avg' in st:
                out[f's{s}/m/{n}'] = st['exp_avg'].numpy().copy()
                out[f's{s}/v/{n}'] = st['exp_avg_sq'].numpy().copy()
    np.savez_compressed(os.path.join(GOLD, f'{name}.npz'), **out)
    print(name, 'ok; loss', [float(out[f's{s}/loss']) for s in range(n_steps)])


def gen_loader_and_eval():
    from hassaku_b200.data.synthetic import write_csv_dataset
    from data.dataset import TrainRecDataset, FullEvalDataset
    from data.dataloader import NegativeSampler, TrainDataLoader
    from torch.utils.data import DataLoader
    from algorithms.sgd_alg import SGDMatrixFactorization
    from eval.eval import FullEvaluator, evaluate_recommender_algorithm
    data = tiny_data()
    with tempfile.TemporaryDirectory() as tmp:
        write_csv_dataset(data, tmp)
        # --- reference loader (data/dataloader.py:92-129) ---
        np.random.seed(64)
        torch.manual_seed(64)
        tds = TrainRecDataset(tmp)
        sampler = NegativeSampler(tds, n_neg=20, neg_sampling_strategy='uniform')
        dl = TrainDataLoader(sampler, tds, batch_size=64, shuffle=True, num_workers=0, prefetch_factor=None)
        out = {}
        for b, (u, i, l) in enumerate(dl):
            if b >= 4:
                break
            out[f'b{b}/u_idxs'], out[f'b{b}/i_idxs'], out[f'b{b}/labels'] = u.numpy(), i.numpy(), l.numpy()
        out['train_indptr'], out['train_indices'] = data.train.indptr, data.train.indices
        np.savez_compressed(os.path.join(GOLD, 'loader_batches.npz'), **out)
        print('loader_batches ok', out['b0/i_idxs'].shape, out['b0/labels'].dtype)

        # --- reference evaluator (eval/eval.py:211-258) for val and test splits ---
        torch.manual_seed(65)
        d = 18
        model = SGDMatrixFactorization(data.n_users, data.n_items, d, use_item_bias=True)
        with torch.no_grad():
            for p in model.parameters():
                p.copy_(torch.randn_like(p) * (1.0 / np.sqrt(d) if p.shape[-1] == d else 0.1))
        ev_out = {f'w/{n}': p.detach().numpy().copy() for n, p in model.state_dict().items()}
        for split in ('val', 'test'):
            eds = FullEvalDataset(tmp, split)
            loader = DataLoader(eds, batch_size=64, num_workers=0)
            evaluator = FullEvaluator(aggr_by_group=True, n_groups=eds.n_user_groups,
                                      user_to_user_group=eds.user_to_user_group)
            res = evaluate_recommender_algorithm(model, loader, evaluator, 'cpu', False)
            ev_out[f'{split}/metric_names'] = np.array(sorted(res.keys()))
            ev_out[f'{split}/metric_values'] = np.array([res[k] for k in sorted(res.keys())], dtype=np.float64)
            # masked scores + top-100 the reference computes inside (eval.py:247-251, :63)
            with torch.no_grad():
                u = torch.arange(data.n_users)
                outm = model.combine_user_item_representations(model.get_user_representations(u),
                                                               model.get_item_representations(
                                                                   torch.arange(data.n_items)))
                mask = torch.tensor(eds.exclude_data[u].A)
                outm[mask] = -torch.inf
                ev_out[f'{split}/masked_scores'] = outm.numpy()
                ev_out[f'{split}/topk_ids'] = outm.topk(k=100).indices.numpy()
            # per-user vectors (aggr_by_group=False path, eval.py:42-46,112-114)
            # (evaluate_recommender_algorithm crashes in log_info_results for vectors, so drive eval_batch directly)
            evaluator2 = FullEvaluator(aggr_by_group=False)
            with torch.no_grad():
                for u_b, _, lab_b in loader:
                    evaluator2.eval_batch(u_b, outm[u_b], lab_b)
            for k, v in evaluator2.get_results().items():
                ev_out[f'{split}/peruser/{k}'] = np.asarray(v)
        ev_out['user_group'] = data.user_group
        np.savez_compressed(os.path.join(GOLD, 'eval_tiny.npz'), **ev_out)
        print('eval_tiny ok', dict(zip(ev_out['val/metric_names'][:3], ev_out['val/metric_values'][:3])))


def gen_metrics_kat():
    from eval.metrics import recall_at_k_batch, precision_at_k_batch, ndcg_at_k_batch
    B, I, k = 10, 20, 10  # framework_tests/eval/test_metrics.py:10-27
    logits = torch.arange(I, 0, -1).repeat(B, 1)
    pats = {'zeros': torch.zeros(B, I), 'ones': torch.ones(B, I)}
    y = torch.zeros(B, I); y[:, 0] = 1; pats['1'] = y
    y = torch.zeros(B, I); y[:, [1, 2]] = 1; pats['2_and_3'] = y
    y = torch.zeros(B, I); y[:, k + 1:] = 1; y[:, 0] = 1; pats['out_of_k'] = y
    out = {'logits': logits.numpy()}
    for n, yt in pats.items():
        out[f'y/{n}'] = yt.numpy()
        out[f'recall/{n}'] = recall_at_k_batch(logits, yt, k=k).item() / B
        out[f'precision/{n}'] = precision_at_k_batch(logits, yt, k=k).item() / B
        out[f'ndcg/{n}'] = ndcg_at_k_batch(logits, yt, k=k).item() / B
    np.savez_compressed(os.path.join(GOLD, 'metrics_kat.npz'), **out)
    print('metrics_kat ok')


def gen_calibration():
    """Reference FullEvaluatorCalibrationDecorator (eval/eval.py:121-208), nested 'tag' + 'pop' like sweep_test.py:68-69,
    over two dense batches; aggregated (with 2 user groups) and per-user (aggr_by_group=False) results."""
    from eval.eval import FullEvaluator, FullEvaluatorCalibrationDecorator
    rng = np.random.RandomState(7)
    U, I, T, P = 40, 150, 7, 3
    logits = torch.from_numpy(rng.randn(U, I).astype(np.float32))
    y_true = torch.from_numpy((rng.rand(U, I) < 0.05).astype(np.float32))
    tags = (rng.rand(I, T) < 0.3).astype(np.float32)
    tags[:5] = 0.                                        # items without tags (data_utils.py:413)
    with np.errstate(invalid='ignore', divide='ignore'):
        item_tag = np.nan_to_num(tags / tags.sum(-1, keepdims=True)).astype(np.float32)
    train = (rng.rand(U, I) < 0.1).astype(np.float32)
    train[:, 0] = 1.                                     # every user has a train item
    ut = train @ item_tag / train.sum(-1, keepdims=True)
    user_tag = (0.01 / T + 0.99 * ut).astype(np.float32)   # eq. 7 smoothing (data_utils.py:425)
    bucket = np.minimum((np.argsort(np.argsort(-train.sum(0))) * P) // I, P - 1)
    item_pop = np.eye(P, dtype=np.float32)[bucket]
    user_pop = (0.01 / P + 0.99 * (train @ item_pop / train.sum(-1, keepdims=True))).astype(np.float32)
    group = rng.randint(0, 2, U)
    out = {'logits': logits.numpy(), 'y_true': y_true.numpy(), 'item_tag': item_tag, 'user_tag': user_tag,
           'item_pop': item_pop, 'user_pop': user_pop, 'user_group': group, 'beta': np.array(0.01)}
    for aggr in (True, False):
        ev = FullEvaluator(aggr_by_group=aggr, n_groups=2, user_to_user_group=torch.from_numpy(group))
        ev = FullEvaluatorCalibrationDecorator(ev, torch.from_numpy(item_tag), torch.from_numpy(user_tag), 'tag', 0.01)
        ev = FullEvaluatorCalibrationDecorator(ev, torch.from_numpy(item_pop), torch.from_numpy(user_pop), 'pop', 0.01)
        for lo, hi in ((0, 24), (24, 40)):
            ev.eval_batch(torch.arange(lo, hi), logits[lo:hi], y_true[lo:hi])
        res = ev.get_results()
        tagk = 'aggr' if aggr else 'peruser'
        out[f'{tagk}/names'] = np.array(sorted(res))
        if aggr:
            out[f'{tagk}/values'] = np.array([float(res[k]) for k in sorted(res)])
        else:
            for k in sorted(res):
                out[f'{tagk}/{k}'] = np.asarray(res[k])
    np.savez_compressed(os.path.join(GOLD, 'calibration_kat.npz'), **out)
    print('calibration_kat ok', len(out['aggr/names']), 'keys')


def gen_calibration_matrices():
    """Reference build_user_and_item_tag_matrix / build_user_and_item_pop_matrix (data/data_utils.py:378-499) on the
    'tiny' synthetic dataset written in the reference's CSV layout plus a random item-tag table."""
    import pandas as pd
    import scipy.sparse as sp
    if not hasattr(sp.csr_matrix, 'A'):     # scipy >= 1.14 dropped `.A`, which data_utils.py:494-499 uses
        sp.csr_matrix.A = property(lambda self: self.toarray())
    from data.data_utils import build_user_and_item_pop_matrix, build_user_and_item_tag_matrix
    from hassaku_b200.data.synthetic import write_csv_dataset
    data = tiny_data()
    T = 9
    rng = np.random.RandomState(1)
    pairs = np.argwhere(rng.rand(data.n_items, T) < 0.25)
    pairs = pairs[pairs[:, 0] >= 4]           # items 0..3 carry no tag
    with tempfile.TemporaryDirectory() as d:
        base = write_csv_dataset(data, os.path.join(d, 'processed_dataset'))
        pd.DataFrame({'tag_idx': np.arange(T)}).to_csv(os.path.join(base, 'tag_idxs.csv'), index=False)
        pd.DataFrame({'item_idx': pairs[:, 0], 'tag_idx': pairs[:, 1]}).to_csv(os.path.join(base, 'item_tag_idxs.csv'),
                                                                              index=False)
        ut, it = build_user_and_item_tag_matrix(d)
        up, ip = build_user_and_item_pop_matrix(d)
    np.savez_compressed(os.path.join(GOLD, 'calibration_matrices.npz'), item_tag_pairs=pairs, n_tags=np.array(T),
                        user_tag=ut.numpy(), item_tag=it.numpy(), user_pop=up.numpy(), item_pop=ip.numpy())
    print('calibration_matrices ok', tuple(ut.shape), tuple(ip.shape), ip.sum(0).tolist())


def gen_baseline_init():
    """Reference SGDBaseline (algorithms/sgd_alg.py:72-107): initial weights under a fixed torch seed (module
    construction + general_weight_init order), parameter names / shapes, and scores of a small batch."""
    from algorithms.sgd_alg import SGDBaseline
    torch.manual_seed(64)
    m = SGDBaseline(37, 23)
    out = {f'init/{n}': p.detach().numpy().copy() for n, p in m.state_dict().items()}
    rng = np.random.RandomState(3)
    u = rng.randint(0, 37, 16).astype(np.int64)
    i = rng.randint(0, 23, (16, 6)).astype(np.int64)
    with torch.no_grad():
        out['u_idxs'], out['i_idxs'] = u, i
        out['scores'] = m(torch.from_numpy(u), torch.from_numpy(i)).numpy()
    np.savez_compressed(os.path.join(GOLD, 'baseline_init.npz'), **out)
    print('baseline_init ok', sorted(k for k in out if k.startswith('init/')))


def main():
    ref_shim.load()
    os.makedirs(GOLD, exist_ok=True)
    torch.set_num_threads(1)
    only = set(sys.argv[1:])      # e.g. `python -m oracle.make_golden calibration baseline`: regenerate just those
    if only:
        if 'calibration' in only:
            gen_calibration()
        if 'calibration_matrices' in only:
            gen_calibration_matrices()
        if 'baseline' in only:
            gen_baseline_init()
            gen_train('train_baseline_bce', 'bce', d=None, use_user_bias=True, use_item_bias=True, use_global_bias=True,
                      optimizer='adamw', B=64, N=4, lr=1e-3, wd=1e-4, model_kind='baseline')
        return
    gen_calibration()
    gen_calibration_matrices()
    gen_baseline_init()
    gen_train('train_baseline_bce', 'bce', d=None, use_user_bias=True, use_item_bias=True, use_global_bias=True,
              optimizer='adamw', B=64, N=4, lr=1e-3, wd=1e-4, model_kind='baseline')
    gen_metrics_kat()
    gen_train('train_bpr', 'bpr', d=18, use_user_bias=False, use_item_bias=True, use_global_bias=False,
              optimizer='adamw', B=96, N=7, lr=3e-4, wd=4e-5)
    gen_train('train_ssm', 'sampled_softmax', d=16, use_user_bias=False, use_item_bias=True, use_global_bias=False,
              optimizer='adamw', B=64, N=12, lr=1e-3, wd=1e-4)
    gen_train('train_bce', 'bce', d=10, use_user_bias=True, use_item_bias=True, use_global_bias=True,
              optimizer='adam', B=64, N=4, lr=1e-3, wd=0.)
    gen_loader_and_eval()


if __name__ == '__main__':
    main()
