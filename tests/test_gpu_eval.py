"""GPU parity tests of the full-rank evaluator (hsk_eval_topk / hsk_topk_merge / hsk_topk_dense / hsk_rank_metrics)
against fixtures generated from the unmodified reference (tests/golden/eval_tiny.npz, metrics_kat.npz) and against the
oracle at BASELINE shapes.

Gates (north_star): masking and top-k item ids bit-exact except documented score ties — ids may differ from the
reference only at ranks where the reference's own fp32 scores of the two items are within 4 ulp (torch.topk's tie order
is implementation-defined and an fp32 GEMM sums in another order than torch's sum(-1)); metrics equal to 1e-6."""
import math

import numpy as np
import pytest
import torch
from scipy import sparse as sp

from hsk_testutil import load_golden

pytestmark = pytest.mark.gpu


def assert_topk_equivalent(ids, ref_ids, ref_scores, ulps=4):
    """ids/ref_ids [B, k]; ref_scores [B, I] = the reference's masked fp32 scores."""
    ids, ref_ids = np.asarray(ids, dtype=np.int64), np.asarray(ref_ids, dtype=np.int64)
    assert ids.shape == ref_ids.shape
    rows, pos = np.nonzero(ids != ref_ids)
    if len(rows) == 0:
        return 0
    a = ref_scores[rows, ids[rows, pos]].astype(np.float64)
    b = ref_scores[rows, ref_ids[rows, pos]].astype(np.float64)
    both_inf = np.isinf(a) & np.isinf(b)
    tol = ulps * np.spacing(np.maximum(np.abs(a), np.abs(b)).astype(np.float32)).astype(np.float64)
    ok = both_inf | (np.abs(a - b) <= tol)
    assert ok.all(), f'{(~ok).sum()} top-k positions differ beyond {ulps} ulp ties'
    return len(rows)


def tiny_data():
    from hassaku_b200.data.synthetic import make_interactions
    return make_interactions(300, 200, 6000, seed=0, n_user_groups=2)


def tiny_model(g):
    from hassaku_b200.algorithms.sgd_alg import SGDMatrixFactorization
    m = SGDMatrixFactorization(300, 200, 18, use_item_bias=True)
    m.load_state_dict({k[2:]: torch.from_numpy(v) for k, v in g.items() if k.startswith('w/')})
    return m.to('cuda')


class _Loader:
    def __init__(self, dataset, batch_size):
        self.dataset, self.batch_size = dataset, batch_size


def make_eval_dataset(data, split):
    from hassaku_b200.data.dataset import FullEvalDataset
    labels = data.val if split == 'val' else data.test
    excl = data.train if split == 'val' else sp.csr_matrix(data.train + data.val)
    return FullEvalDataset.from_interactions(labels, excl, split, data.user_group, data.n_user_groups)


@pytest.mark.parametrize('split', ['val', 'test'])
@pytest.mark.parametrize('batch_size', [64, 300])
def test_evaluate_matches_reference_fixture(split, batch_size):
    from hassaku_b200.eval.eval import FullEvaluator, evaluate_recommender_algorithm, TopKScorer, device_csr
    g = load_golden('eval_tiny')
    data = tiny_data()
    model = tiny_model(g)
    ds = make_eval_dataset(data, split)
    ev = FullEvaluator(aggr_by_group=True, n_groups=ds.n_user_groups, user_to_user_group=ds.user_to_user_group)
    res = evaluate_recommender_algorithm(model, _Loader(ds, batch_size), ev, 'cuda')
    names = [str(x) for x in g[f'{split}/metric_names']]
    assert sorted(res.keys()) == names
    for n, v in zip(names, g[f'{split}/metric_values']):
        assert abs(res[n] - v) <= 1e-6, (n, res[n], v)
    # key order of the reference dict: ALL group first, k descending, precision/recall/ndcg
    assert list(res.keys())[:4] == ['precision@100', 'recall@100', 'ndcg@100', 'precision@50']
    # ids
    scorer = TopKScorer(model, 300, 100)
    u = torch.arange(300, device='cuda')
    scores, ids = scorer(u, device_csr(ds, 'exclude_data', 'cuda'))
    ref_scores = g[f'{split}/masked_scores']
    assert_topk_equivalent(ids.cpu().numpy(), g[f'{split}/topk_ids'], ref_scores)
    got = np.take_along_axis(ref_scores, ids.cpu().numpy().astype(np.int64), axis=1)
    np.testing.assert_allclose(scores.cpu().numpy(), got, rtol=1e-5, atol=1e-6)
    # masking: no excluded item may appear ahead of a finite score
    excl = ds.exclude_data
    ids_np = ids.cpu().numpy()
    for r in range(300):
        row = excl.indices[excl.indptr[r]:excl.indptr[r + 1]]
        assert not np.isin(ids_np[r][np.isfinite(scores[r].cpu().numpy())], row).any()


def test_per_user_vectors_match_reference_fixture():
    from hassaku_b200.eval.eval import FullEvaluator, evaluate_recommender_algorithm
    g = load_golden('eval_tiny')
    data = tiny_data()
    model = tiny_model(g)
    ds = make_eval_dataset(data, 'val')
    res = evaluate_recommender_algorithm(model, _Loader(ds, 128), FullEvaluator(aggr_by_group=False), 'cuda')
    for k in (5, 10, 50, 100):
        for m in ('precision', 'recall', 'ndcg'):
            np.testing.assert_allclose(res[f'{m}@{k}'], g[f'val/peruser/{m}@{k}'], rtol=0, atol=1e-6)


def test_dense_eval_batch_api_matches_reference_fixture():
    """FullEvaluator.eval_batch(u_idxs, logits, y_true) with dense matrices (the reference's own calling convention)."""
    from hassaku_b200.eval.eval import FullEvaluator
    g = load_golden('eval_tiny')
    data = tiny_data()
    ev = FullEvaluator(aggr_by_group=True, n_groups=2, user_to_user_group=torch.Tensor(data.user_group))
    logits = torch.from_numpy(g['val/masked_scores']).cuda()
    y = torch.from_numpy(data.val.toarray().astype('float32')).cuda()
    for s in range(0, 300, 64):
        ev.eval_batch(torch.arange(s, min(s + 64, 300)), logits[s:s + 64], y[s:s + 64])
    res = ev.get_results()
    for n, v in zip([str(x) for x in g['val/metric_names']], g['val/metric_values']):
        assert abs(res[n] - v) <= 1e-6, n


def test_metric_known_answers_of_reference_tests():
    """framework_tests/eval/test_metrics.py:29-69 through the CUDA metric functions."""
    from hassaku_b200.eval.metrics import recall_at_k_batch, precision_at_k_batch, ndcg_at_k_batch
    B, I, k = 10, 20, 10
    logits = torch.arange(I, 0, -1).repeat(B, 1).float().cuda()
    y0, y1 = torch.zeros(B, I).cuda(), torch.ones(B, I).cuda()
    ya = torch.zeros(B, I); ya[:, 0] = 1
    yb = torch.zeros(B, I); yb[:, [1, 2]] = 1
    yc = torch.zeros(B, I); yc[:, k + 1:] = 1; yc[:, 0] = 1
    ya, yb, yc = ya.cuda(), yb.cuda(), yc.cuda()
    mean = lambda f, y: f(logits, y, k=k).item() / B
    assert mean(recall_at_k_batch, y0) == 0
    assert mean(recall_at_k_batch, y1) == pytest.approx(k / I, abs=1e-6)
    assert mean(recall_at_k_batch, ya) == 1
    assert mean(recall_at_k_batch, yb) == 1
    assert mean(recall_at_k_batch, yc) == pytest.approx(1 / (I - k), abs=1e-6)
    assert mean(precision_at_k_batch, y0) == 0
    assert mean(precision_at_k_batch, y1) == 1
    assert mean(precision_at_k_batch, ya) == pytest.approx(1 / k, abs=1e-6)
    assert mean(precision_at_k_batch, yb) == pytest.approx(2 / k, abs=1e-6)
    assert mean(precision_at_k_batch, yc) == pytest.approx(1 / k, abs=1e-6)
    disc = 1. / torch.log2(torch.arange(2, k + 2).float())
    assert mean(ndcg_at_k_batch, y0) == 0
    assert mean(ndcg_at_k_batch, y1) == pytest.approx(1, abs=1e-6)
    assert mean(ndcg_at_k_batch, ya) == pytest.approx(1, abs=1e-6)
    assert mean(ndcg_at_k_batch, yb) == pytest.approx(
        (math.log2(4) + math.log2(3)) / (math.log2(4) * (1 + math.log2(3))), abs=1e-5)
    assert mean(ndcg_at_k_batch, yc) == pytest.approx(1 / disc[:min(k, I - k)].sum().item(), abs=1e-6)
    # fixture computed by the reference's own functions
    g = load_golden('metrics_kat')
    for n in ('zeros', 'ones', '1', '2_and_3', 'out_of_k'):
        y = torch.from_numpy(g[f'y/{n}']).cuda()
        assert recall_at_k_batch(logits, y, k=k).item() / B == pytest.approx(float(g[f'recall/{n}']), abs=1e-6)
        assert precision_at_k_batch(logits, y, k=k).item() / B == pytest.approx(float(g[f'precision/{n}']), abs=1e-6)
        assert ndcg_at_k_batch(logits, y, k=k).item() / B == pytest.approx(float(g[f'ndcg/{n}']), abs=1e-6)
    with pytest.raises(AssertionError):
        recall_at_k_batch(logits, y0, k=5, idx_topk=torch.zeros(B, 10, dtype=torch.int64))
    assert precision_at_k_batch(logits, ya, k=k, aggr_sum=False).shape == (B,)


@pytest.mark.parametrize('U,I,d,n_eval,biases', [
    (6040, 3706, 402, 1024, (False, True, False)),    # cfg1/cfg2 shape, user subsample sized for the CPU oracle
    (3000, 10677, 128, 512, (False, True, False)),    # cfg3 item count
    (500, 1000, 64, 500, (True, True, True)),
    (130, 129, 5, 130, (False, False, False)),        # ragged: partial tiles everywhere
])
def test_eval_vs_oracle_at_baseline_shapes(U, I, d, n_eval, biases):
    from oracle import mf_oracle as O
    from hassaku_b200.algorithms.sgd_alg import SGDMatrixFactorization
    from hassaku_b200.data.dataset import FullEvalDataset
    from hassaku_b200.data.synthetic import make_interactions
    from hassaku_b200.eval.eval import FullEvaluator, TopKScorer, device_csr
    torch.manual_seed(11)
    data = make_interactions(U, I, min(U * I // 8, U * 150), seed=3, n_user_groups=2)
    ref = O.OracleMF(U, I, d, *biases)
    with torch.no_grad():
        for p in ref.parameters():
            p.copy_(torch.randn_like(p) * (1.0 / math.sqrt(d) if p.shape[-1] == d else 0.1))
    model = SGDMatrixFactorization(U, I, d, *biases)
    model.load_state_dict(ref.state_dict())
    model.to('cuda')
    users = np.sort(np.random.RandomState(0).choice(U, n_eval, replace=False))
    res_ref, topk_ref = O.evaluate(ref, data.val, data.train, 256, 2, torch.from_numpy(data.user_group).float(),
                                   users=users, return_topk=True)
    ds = FullEvalDataset.from_interactions(data.val, data.train, 'val', data.user_group, 2)
    ev = FullEvaluator(True, 2, ds.user_to_user_group)
    scorer = TopKScorer(model, 256, 100)
    labels, excl = device_csr(ds, 'iteration_matrix', 'cuda'), device_csr(ds, 'exclude_data', 'cuda')
    ids_all = []
    for s in range(0, n_eval, 256):
        u = torch.from_numpy(users[s:s + 256].astype(np.int64)).cuda()
        _, ids = scorer(u, excl)
        ev.eval_batch_topk(u, ids, labels)
        ids_all.append(ids.cpu().numpy().copy())
    res = ev.get_results()
    for n, v in res_ref.items():
        assert abs(res[n] - v) <= 1e-6, (n, res[n], v)
    ref_scores = O.masked_scores(ref, torch.from_numpy(users.astype(np.int64)), data.train).numpy()
    n_diff = assert_topk_equivalent(np.concatenate(ids_all), topk_ref.numpy(), ref_scores)
    assert n_diff <= 0.001 * n_eval * 100  # position-wise agreement >= 99.9 %


def test_split_plan_and_merge_are_order_independent():
    """Small user batches split the item range over several CTAs and merge; the result must equal the one-split run."""
    from hassaku_b200 import _C
    from hassaku_b200.algorithms.sgd_alg import SGDMatrixFactorization
    torch.manual_seed(5)
    U, I, d, k = 64, 20000, 32, 100
    model = SGDMatrixFactorization(U, I, d, use_item_bias=True).to('cuda')
    with torch.no_grad():
        model.item_embeddings.weight.mul_(300.)
        model.user_embeddings.weight.mul_(300.)
    u = torch.arange(U, device='cuda')
    scores = model(u.repeat_interleave(1), torch.arange(I, device='cuda').repeat(U, 1))
    ref_s, ref_i = scores.topk(k, dim=1)
    out_s = torch.empty((U, k), device='cuda')
    out_i = torch.empty((U, k), dtype=torch.int32, device='cuda')
    scratch = torch.empty(_C.eval_topk_scratch_bytes(U, I, k), dtype=torch.uint8, device='cuda')
    assert scratch.numel() > U * 256 * 8  # more than one split planned
    _C.eval_topk(model._tables(), u, k, out_s, out_i, scratch)
    assert_topk_equivalent(out_i.cpu().numpy(), ref_i.cpu().numpy(), scores.detach().cpu().numpy())
    # hsk_topk_merge: shard the item table by (i mod G), score each shard, merge
    G = 4
    parts_s, parts_i = [], []
    Vw, Ib = model.item_embeddings.weight.detach(), model.item_bias.weight.detach()
    for r in range(G):
        Vr = Vw[r::G].contiguous()
        Ir = Ib[r::G].contiguous()
        t = _C.make_tables(model.user_embeddings.weight.detach(), Vr, None, Ir, None, d)
        ps = torch.empty((U, k), device='cuda')
        pi = torch.empty((U, k), dtype=torch.int32, device='cuda')
        sc = torch.empty(_C.eval_topk_scratch_bytes(U, Vr.shape[0], k), dtype=torch.uint8, device='cuda')
        _C.eval_topk(t, u, k, ps, pi, sc, id_offset=r, id_stride=G)
        parts_s.append(ps)
        parts_i.append(pi)
    ms = torch.empty((U, k), device='cuda')
    mi = torch.empty((U, k), dtype=torch.int32, device='cuda')
    _C.topk_merge(torch.stack(parts_s), torch.stack(parts_i), ms, mi)
    assert torch.equal(mi, out_i) and torch.equal(ms, out_s)


def test_ties_resolve_to_lower_item_id_and_masked_items_rank_last():
    from hassaku_b200 import _C
    from hassaku_b200.algorithms.sgd_alg import SGDMatrixFactorization
    U, I, d, k = 3, 300, 8, 100
    model = SGDMatrixFactorization(U, I, d).to('cuda')
    with torch.no_grad():
        model.item_embeddings.weight.zero_()          # every score ties at 0
        model.item_embeddings.weight[250:260] = 1.0   # ten equal winners for a positive user
        model.user_embeddings.weight.fill_(1.0)
    u = torch.arange(U, device='cuda')
    # user 0: nothing excluded; user 1: the winners excluded; user 2: all but 50 items excluded
    ex = sp.lil_matrix((U, I), dtype=bool)
    ex[1, 250:260] = True
    ex[2, 50:] = True
    ex = sp.csr_matrix(ex)
    indptr = torch.from_numpy(ex.indptr.astype(np.int64)).cuda()
    indices = torch.from_numpy(ex.indices.astype(np.int32)).cuda()
    out_s = torch.empty((U, k), device='cuda')
    out_i = torch.empty((U, k), dtype=torch.int32, device='cuda')
    scratch = torch.empty(_C.eval_topk_scratch_bytes(U, I, k), dtype=torch.uint8, device='cuda')
    _C.eval_topk(model._tables(), u, k, out_s, out_i, scratch, indptr, indices)
    i0, i1, i2 = out_i.cpu().numpy()
    assert list(i0[:10]) == list(range(250, 260)) and list(i0[10:]) == list(range(0, 90))
    assert list(i1) == list(range(0, 100))
    assert list(i2[:50]) == list(range(0, 50)) and list(i2[50:]) == list(range(50, 100))  # -inf ties by id
    s2 = out_s.cpu().numpy()[2]
    assert np.isfinite(s2[:50]).all() and np.isneginf(s2[50:]).all()


def test_too_few_items_raises_like_reference():
    from hassaku_b200.algorithms.sgd_alg import SGDMatrixFactorization
    from hassaku_b200.data.dataset import FullEvalDataset
    from hassaku_b200.eval.eval import FullEvaluator, evaluate_recommender_algorithm
    model = SGDMatrixFactorization(10, 50, 8).to('cuda')
    z = sp.csr_matrix((10, 50), dtype=np.int16)
    ds = FullEvalDataset.from_interactions(z, z)
    with pytest.raises(RuntimeError):
        evaluate_recommender_algorithm(model, _Loader(ds, 8), FullEvaluator(), 'cuda')
