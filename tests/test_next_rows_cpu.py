"""CPU: the SURVEY §8(f) rows that are host logic over the kernels' outputs — calibration metrics / decorator (row 3) and
the SGDBaseline model shell (row 4) — against fixtures generated from the unmodified reference
(oracle/make_golden.py: calibration_kat.npz, baseline_init.npz)."""
import numpy as np
import pytest
import torch

from hsk_testutil import load_golden


class _StubEvaluator:
    """Stands in for hassaku_b200.eval.eval.FullEvaluator (whose metric kernels need a GPU): the decorator only needs
    the group bookkeeping and something to delegate to."""
    K_VALUES = [5, 10, 50, 100]

    def __init__(self, aggr_by_group, n_groups, user_to_user_group):
        self.aggr_by_group, self.n_groups, self.user_to_user_group = aggr_by_group, n_groups, user_to_user_group
        self.calls = 0

    def _reset_internal_dict(self):
        self.calls = 0

    def get_n_groups(self):
        return self.n_groups

    def get_user_to_user_group(self):
        return self.user_to_user_group

    def eval_batch(self, u, logits, y):
        self.calls += 1

    def eval_batch_topk(self, u, ids, labels):
        self.calls += 1

    def get_results(self):
        out = {'inner': float(self.calls)}
        self._reset_internal_dict()
        return out


def _decorated(g, aggr):
    from hassaku_b200.eval.eval import FullEvaluatorCalibrationDecorator as Deco
    ev = _StubEvaluator(aggr, 2, torch.from_numpy(g['user_group']))
    ev = Deco(ev, torch.from_numpy(g['item_tag']), torch.from_numpy(g['user_tag']), 'tag', float(g['beta']))
    return Deco(ev, torch.from_numpy(g['item_pop']), torch.from_numpy(g['user_pop']), 'pop', float(g['beta']))


@pytest.mark.parametrize('path', ['dense', 'topk'])
def test_calibration_decorator_aggregated_vs_reference(path):
    g = load_golden('calibration_kat')
    ev = _decorated(g, True)
    logits, y = torch.from_numpy(g['logits']), torch.from_numpy(g['y_true'])
    for lo, hi in ((0, 24), (24, 40)):
        u = torch.arange(lo, hi)
        if path == 'dense':
            ev.eval_batch(u, logits[lo:hi], y[lo:hi])
        else:   # what the fused evaluator hands over: ranked top-100 ids
            ev.eval_batch_topk(u, logits[lo:hi].topk(100).indices.int(), None)
    res = ev.get_results()
    assert res.pop('inner') == 2.0                       # both decorators delegated to the wrapped evaluator once per batch
    want = {k: v for k, v in zip(g['aggr/names'], g['aggr/values']) if 'tag_' in k or 'pop_' in k}
    assert sorted(res) == sorted(want) and len(want) == 2 * 3 * 4 * 3   # 2 prefixes x 3 distances x 4 k x (ALL + 2 groups)
    for k, v in want.items():
        assert np.isclose(res[k], v, rtol=1e-5, atol=1e-7, equal_nan=True), (k, res[k], v)
    assert ev.get_results() == {'inner': 0.0}            # reset after get_results (eval.py:116)


def test_calibration_decorator_per_user_vectors_vs_reference():
    g = load_golden('calibration_kat')
    ev = _decorated(g, False)
    logits, y = torch.from_numpy(g['logits']), torch.from_numpy(g['y_true'])
    for lo, hi in ((0, 24), (24, 40)):
        ev.eval_batch(torch.arange(lo, hi), logits[lo:hi], y[lo:hi])
    res = ev.get_results()
    n = 0
    for k in g['peruser/names']:
        if 'tag_' in k or 'pop_' in k:
            want = g[f'peruser/{k}']
            assert res[k].shape == want.shape, k
            assert np.allclose(res[k], want, rtol=1e-5, atol=1e-7, equal_nan=True), k
            n += 1
    assert n == 72


def test_calibration_distances_closed_forms():
    from hassaku_b200.eval.metrics import hellinger_distance, jensen_shannon_distance, kl_divergence
    p = torch.tensor([[0.5, 0.5, 0.0], [0.2, 0.3, 0.5]])
    assert torch.allclose(hellinger_distance(p, p), torch.zeros(2))
    one = torch.tensor([[1.0, 0.0]]); other = torch.tensor([[0.0, 1.0]])
    assert torch.allclose(hellinger_distance(one, other), torch.ones(1))          # disjoint supports: distance 1
    q = torch.tensor([[0.25, 0.25, 0.5], [0.2, 0.3, 0.5]])
    assert abs(float(kl_divergence(q[1:], q[1:]))) < 1e-7
    a, b = torch.tensor([[0.5, 0.5]]), torch.tensor([[0.25, 0.75]])
    want = 0.5 * np.log(0.5 / 0.25) + 0.5 * np.log(0.5 / 0.75)
    assert abs(float(kl_divergence(a, b)) - want) < 1e-6
    assert torch.allclose(jensen_shannon_distance(a, b), jensen_shannon_distance(b, a))   # symmetric
    assert torch.isnan(kl_divergence(p[:1], q[:1])).all()                                  # 0 * log 0 -> nan, like the reference


def test_calibration_decorator_rejects_bad_beta_and_short_topk():
    from hassaku_b200.eval.eval import FullEvaluatorCalibrationDecorator as Deco
    g = load_golden('calibration_kat')
    with pytest.raises(AssertionError):
        Deco(_StubEvaluator(True, 0, None), torch.from_numpy(g['item_tag']), torch.from_numpy(g['user_tag']), 'tag', 1.5)
    ev = Deco(_StubEvaluator(True, 0, None), torch.from_numpy(g['item_tag']), torch.from_numpy(g['user_tag']))
    with pytest.raises(AssertionError):
        ev.eval_batch_topk(torch.arange(4), torch.zeros((4, 50), dtype=torch.int32), None)


def test_sgd_baseline_shell_matches_reference_init_names_and_shapes():
    from hassaku_b200 import _C
    from hassaku_b200.algorithms.sgd_alg import SGDBaseline, SGDMatrixFactorization
    g = load_golden('baseline_init')
    torch.manual_seed(64)
    m = SGDBaseline(37, 23)
    sd = m.state_dict()
    assert sorted(sd) == ['global_bias', 'item_bias.weight', 'user_bias.weight']     # base_classes.py:156-165 round trip
    for k, v in sd.items():
        assert np.array_equal(v.numpy(), g['init/' + k]), k                          # same seed -> bit-identical init
    assert [n for n, _ in m.named_parameters()] == ['global_bias', 'user_bias.weight', 'item_bias.weight']
    assert isinstance(m, SGDMatrixFactorization) and m.name == 'SGDBaseline' and m.embedding_dim == 1
    assert float(m.user_embeddings.weight.abs().max()) == 0.0 and float(m.item_embeddings.weight.abs().max()) == 0.0
    # materialising accessors follow the reference (sgd_alg.py:93-102)
    u = torch.from_numpy(g['u_idxs']); i = torch.from_numpy(g['i_idxs'])
    out = m.combine_user_item_representations(m.get_user_representations(u), m.get_item_representations(i))
    assert np.allclose(out.detach().numpy(), g['scores'], rtol=1e-6, atol=1e-7)
    with pytest.raises(_C.HskError):
        m(u, i)                                                                      # no CPU path
    m2 = SGDBaseline.build_from_conf({}, type('D', (), {'n_users': 37, 'n_items': 23})())
    m2.load_state_dict(sd)
    assert torch.equal(m2.arena, m.arena)


def test_factor_model_detection_builds_a_frozen_mf_shell():
    """SVD / ALS (`users_factors`, `items_factors`) and RBMF (`X`, `C`) are scored by the MF kernels."""
    from hassaku_b200.algorithms.sgd_alg import SGDMatrixFactorization
    from hassaku_b200.eval.eval import factor_model_of
    rng = np.random.RandomState(0)

    class SVDLike:
        name = 'SVDAlgorithm'
        users_factors = rng.randn(11, 6)            # float64, as scipy's svds returns
        items_factors = rng.randn(7, 6)

    class RBMFLike:
        X = rng.randn(5, 3).astype(np.float32)
        C = rng.randn(9, 3).astype(np.float32)

    a = SVDLike()
    m = factor_model_of(a, device='cpu')
    assert isinstance(m, SGDMatrixFactorization) and (m.n_users, m.n_items, m.embedding_dim) == (11, 7, 6)
    assert not (m.use_user_bias or m.use_item_bias or m.use_global_bias) and 'SVDAlgorithm' in m.name
    assert np.allclose(m.user_embeddings.weight.detach().numpy(), a.users_factors.astype(np.float32))
    assert np.allclose(m.item_embeddings.weight.detach().numpy(), a.items_factors.astype(np.float32))
    assert factor_model_of(a, device='cpu') is m                       # cached while the factor arrays are the same objects
    a.users_factors = a.users_factors.copy()                           # a new fit -> rebuilt
    assert factor_model_of(a, device='cpu') is not m
    r = factor_model_of(RBMFLike(), device='cpu')
    assert (r.n_users, r.n_items, r.embedding_dim) == (5, 9, 3)
    assert factor_model_of(object()) is None

    class Unfitted:
        users_factors = None
        items_factors = None

    assert factor_model_of(Unfitted()) is None

    class Mismatch:
        users_factors = np.zeros((4, 3))
        items_factors = np.zeros((4, 2))

    assert factor_model_of(Mismatch()) is None


def test_calibration_matrices_vs_reference(tmp_path):
    """build_user_and_item_{tag,pop}_matrix on the tiny dataset: in-memory and CSV entry points against the reference's
    outputs (oracle/make_golden.py gen_calibration_matrices)."""
    import pandas as pd
    from hassaku_b200.data import calibration as C
    from hassaku_b200.data.synthetic import make_interactions, write_csv_dataset
    g = load_golden('calibration_matrices')
    data = make_interactions(300, 200, 6000, seed=0, n_user_groups=2)
    pairs, T = g['item_tag_pairs'], int(g['n_tags'])
    ut, it = C.tag_matrices_from_csr(data.train.tocsr(), pairs[:, 0], pairs[:, 1], T)
    up, ip = C.pop_matrices_from_csr(data.train.tocsr())
    assert np.array_equal(it.numpy(), g['item_tag']) and np.array_equal(ip.numpy(), g['item_pop'])
    assert np.allclose(ut.numpy(), g['user_tag'], rtol=1e-6, atol=1e-8, equal_nan=True)
    assert np.allclose(up.numpy(), g['user_pop'], rtol=1e-6, atol=1e-8, equal_nan=True)
    assert float(it[:4].abs().max()) == 0.0                                   # untagged items: zero rows
    assert ut.dtype == torch.float32 and ip.shape == (200, 3) and float(ip.sum()) == 200.0
    base = write_csv_dataset(data, str(tmp_path / 'processed_dataset'))
    pd.DataFrame({'tag_idx': np.arange(T)}).to_csv(base + '/tag_idxs.csv', index=False)
    pd.DataFrame({'item_idx': pairs[:, 0], 'tag_idx': pairs[:, 1]}).to_csv(base + '/item_tag_idxs.csv', index=False)
    ut2, it2 = C.build_user_and_item_tag_matrix(str(tmp_path))
    up2, ip2 = C.build_user_and_item_pop_matrix(str(tmp_path))
    assert torch.equal(it2, it) and torch.equal(ip2, ip)
    assert torch.allclose(ut2, ut, equal_nan=True) and torch.allclose(up2, up, equal_nan=True)
    with pytest.raises(AssertionError):
        C.pop_matrices_from_csr(data.train.tocsr(), alpha_smoothening=2.0)


def test_hit_accumulation_and_keys():
    """Hit@k (north_star; not in the reference): derived from the precision column, per group, opt-in."""
    from hassaku_b200.eval.eval import FullEvaluator, accumulate_hits
    from hassaku_b200.eval.metrics import hit_from_precision
    pu = torch.tensor([[[0.2, 0, 0], [0, 0, 0]], [[0, 0, 0], [0, 0, 0]], [[0.1, 0, 0], [0.5, 0, 0]]])
    assert hit_from_precision(pu).tolist() == [[1, 0], [0, 0], [1, 1]]
    h = accumulate_hits(None, pu, torch.tensor([0, 1, 1]), 2)
    assert h.tolist() == [[2, 1], [1, 0], [1, 1]]
    h = accumulate_hits(h, pu, torch.tensor([1, 1, 0]), 2)
    assert h.tolist() == [[4, 2], [2, 1], [2, 1]]
    assert accumulate_hits(None, pu, None, 0).tolist() == [[2, 1]]
    assert FullEvaluator(True, 0, None).hit is False            # default keeps the reference's 12 (1 + n_groups) keys
    ev = FullEvaluator(True, 2, torch.tensor([0, 1, 1]), hit=True)
    assert ev._per_user_buf(3, 'cpu') is not None and FullEvaluator(True, 0, None)._per_user_buf(3, 'cpu') is None


@pytest.mark.parametrize('name', ['ACF', 'UProtoMF', 'IProtoMF', 'UIProtoMF'])
def test_dot_product_sgd_models_reduce_to_mf_factors(name):
    """The reference's prototype / anchor models score by a dot product of representations (sgd_alg.py:262-267, 353-356,
    450-453, 535-543): the frozen MF shell built from the representation matrices reproduces `predict` exactly."""
    from oracle import ref_shim
    if not ref_shim.available():
        pytest.skip('reference checkout not mounted')
    ref_shim.load()
    import algorithms.sgd_alg as R
    from hassaku_b200.eval.eval import factor_model_of
    torch.manual_seed(0)
    U, I = 23, 31
    model = {'ACF': lambda: R.ACF(U, I, embedding_dim=12, n_anchors=5),
             'UProtoMF': lambda: R.UProtoMF(U, I, embedding_dim=12, n_prototypes=7),
             'IProtoMF': lambda: R.IProtoMF(U, I, embedding_dim=12, n_prototypes=7),
             'UIProtoMF': lambda: R.UIProtoMF(U, I, embedding_dim=12, u_n_prototypes=7, i_n_prototypes=5)}[name]()
    with torch.no_grad():
        for p in model.parameters():
            p.copy_(torch.randn_like(p))
    shell = factor_model_of(model, device='cpu')
    assert shell is not None and (shell.n_users, shell.n_items) == (U, I)
    assert model.training                                             # mode restored (the reference's predict() leaves eval())
    u = torch.arange(U)
    i = torch.arange(I).repeat(U, 1)
    want = model.predict(u, i)                                        # base_classes.py:150-154, the evaluator's call
    got = shell.user_embeddings.weight.detach() @ shell.item_embeddings.weight.detach().T
    assert torch.allclose(got, want, rtol=1e-5, atol=1e-6), float((got - want).abs().max())


@pytest.mark.parametrize('model_kind', ['mf_item_bias', 'mf_all_biases', 'baseline'])
def test_autograd_plumbing_of_the_score_function_with_stubbed_kernels(model_kind, monkeypatch):
    """Host logic only: `_MfScoreFn` routes the dense gradient tables of hsk_mf_scatter_grads to exactly the parameters
    that exist (None for absent tables, SGDBaseline has no embedding parameters).  The two kernels are replaced by torch
    stand-ins here — the kernels themselves are tested on the GPU."""
    from hassaku_b200 import _C
    from hassaku_b200.algorithms import sgd_alg as A
    U, I, d, B, N1 = 13, 11, 5, 7, 4
    torch.manual_seed(0)
    if model_kind == 'baseline':
        model = A.SGDBaseline(U, I)
    else:
        model = A.SGDMatrixFactorization(U, I, d, model_kind == 'mf_all_biases', True, model_kind == 'mf_all_biases')
    with torch.no_grad():
        for p in model.parameters():
            p.copy_(torch.randn_like(p))
    lay = model.layout

    def dense_scores(arena, u, i):
        Uw, Vw, Ub, Ib, Gb = lay.views(arena)
        s = (Uw[u][:, None, :] * Vw[i]).sum(-1)
        if Ub is not None:
            s = s + Ub[u]
        if Ib is not None:
            s = s + Ib[i].squeeze(-1)
        if Gb is not None:
            s = s + Gb
        return s

    def fake_scores(tables, u, i, out, status=None):
        out.copy_(dense_scores(model.arena.detach(), u, i))

    def fake_scatter(tables, gtables, u, i, ds, status=None):
        with torch.enable_grad():                  # Function.backward runs with grad mode off
            a = model.arena.detach().clone().requires_grad_()
            (dense_scores(a, u, i) * ds).sum().backward()
        fake_scatter.g_arena.copy_(a.grad)

    orig_tables = lay.tables

    def tables_spy(arena):
        if arena is not model.arena and arena.numel() == lay.n_total and float(arena.abs().sum()) == 0.0:
            fake_scatter.g_arena = arena           # the fresh gradient arena backward() allocates
        return orig_tables(arena)

    monkeypatch.setattr(_C, 'make_tables', lambda *args, **kw: None)   # the real one (rightly) refuses host tensors
    monkeypatch.setattr(_C, 'mf_scores', fake_scores)
    monkeypatch.setattr(_C, 'mf_scatter_grads', fake_scatter)
    monkeypatch.setattr(lay, 'tables', tables_spy)
    u = torch.randint(0, U, (B,))
    i = torch.randint(0, I, (B, N1))
    Ub = model.user_bias.weight if model.use_user_bias else None
    Ib = model.item_bias.weight if model.use_item_bias else None
    Gb = model.global_bias if model.use_global_bias else None
    if model_kind == 'baseline':
        out = A._MfScoreFn.apply(model, u, i, None, None, Ub, Ib, Gb)
    else:
        out = A._MfScoreFn.apply(model, u, i, model.user_embeddings.weight, model.item_embeddings.weight, Ub, Ib, Gb)
    w = torch.randn_like(out)
    (out * w).sum().backward()
    # reference: plain autograd through the dense formula on a copy of the arena
    a = model.arena.detach().clone().requires_grad_()
    (dense_scores(a, u, i) * w).sum().backward()
    want = dict(zip(['user_embeddings.weight', 'item_embeddings.weight', 'user_bias.weight', 'item_bias.weight', 'global_bias'],
                    lay.views(a.grad)))
    names = [n for n, _ in model.named_parameters()]
    assert ('user_embeddings.weight' in names) == (model_kind != 'baseline')
    for n, p in model.named_parameters():
        assert p.grad is not None, n
        assert torch.allclose(p.grad, want[n].reshape(p.grad.shape), rtol=1e-6, atol=1e-7), n
