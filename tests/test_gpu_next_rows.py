"""GPU: SURVEY §8(f) rows 3-4 on the real kernels — SGDBaseline through the score / fused-step / evaluator kernels and
the calibration decorator over the device evaluator — plus hsk_shard_local_index and the ring-vs-register kernel A/B.
The host logic they cover is also tested on the CPU in tests/test_next_rows_cpu.py against the same reference fixtures."""
import numpy as np
import pytest
import torch

from hsk_testutil import load_golden

pytestmark = [pytest.mark.gpu]
RTOL = 1e-5
BIAS_NAMES = ['user_bias.weight', 'item_bias.weight', 'global_bias']


def rel_err(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))


def _baseline(g, prefix='init/'):
    from hassaku_b200.algorithms.sgd_alg import SGDBaseline
    U, I, _, B, N = [int(x) for x in g['meta_dims']]
    m = SGDBaseline(U, I)
    m.load_state_dict({k[len(prefix):]: torch.from_numpy(v) for k, v in g.items() if k.startswith(prefix)})
    return m.to('cuda'), (U, I, B, N)


def test_baseline_forward_loss_backward_vs_reference_fixture():
    from hassaku_b200.train.rec_losses import RecBinaryCrossEntropy
    g = load_golden('train_baseline_bce')
    model, (U, I, B, N) = _baseline(g)
    loss_fn = RecBinaryCrossEntropy()
    for s in range(3):
        if s > 0:
            with torch.no_grad():
                for n, p in model.named_parameters():
                    p.copy_(torch.from_numpy(g[f's{s - 1}/param/{n}']).cuda())
        u = torch.from_numpy(g[f's{s}/u_idxs']).cuda()
        i = torch.from_numpy(g[f's{s}/i_idxs']).cuda()
        labels = torch.zeros(i.shape, dtype=torch.float64, device='cuda')
        labels[:, 0] = 1.
        model.zero_grad()
        out = model(u, i)
        assert rel_err(out.detach().cpu().numpy(), g[f's{s}/scores']) < RTOL
        loss = loss_fn.compute_loss(out, labels)
        assert abs(loss.item() - float(g[f's{s}/loss'])) <= RTOL * abs(float(g[f's{s}/loss']))
        loss.backward()
        for n, p in model.named_parameters():
            assert rel_err(p.grad.cpu().numpy(), g[f's{s}/grad/{n}']) < RTOL, n
        model.check_status()


def test_baseline_fused_step_teacher_forced_vs_reference_fixture():
    from hassaku_b200.train.optim import DenseAdam
    from hassaku_b200.train.rec_losses import RecBinaryCrossEntropy
    from hassaku_b200.train.trainer_step import FusedMFTrainStep
    g = load_golden('train_baseline_bce')
    model, (U, I, B, N) = _baseline(g)
    lr, wd = [float(x) for x in g['meta_hparams']]
    opt = DenseAdam(model, lr=lr, weight_decay=wd, decoupled=True, arith=1)
    step = FusedMFTrainStep(model, RecBinaryCrossEntropy(), opt)
    views = lambda arena: dict(zip(BIAS_NAMES, model.layout.views(arena)[2:]))
    for s in range(3):
        if s > 0:
            with torch.no_grad():
                for n, p in model.named_parameters():
                    p.copy_(torch.from_numpy(g[f's{s - 1}/param/{n}']).cuda())
            for n in BIAS_NAMES:
                views(opt.m)[n].copy_(torch.from_numpy(g[f's{s - 1}/m/{n}']).cuda().view_as(views(opt.m)[n]))
                views(opt.v)[n].copy_(torch.from_numpy(g[f's{s - 1}/v/{n}']).cuda().view_as(views(opt.v)[n]))
        opt.t = s
        step(torch.from_numpy(g[f's{s}/u_idxs']), torch.from_numpy(g[f's{s}/i_idxs']))
        loss = step.pop_loss_sum()
        assert abs(loss - float(g[f's{s}/loss'])) <= RTOL * abs(float(g[f's{s}/loss']))
        for n, p in model.named_parameters():
            assert rel_err(p.detach().cpu().numpy(), g[f's{s}/param/{n}']) < RTOL + 2e-3 * lr, n
        # the all-zero embedding tables never move
        assert float(model.user_embeddings.weight.abs().max()) == 0.0
        assert float(model.item_embeddings.weight.abs().max()) == 0.0


def test_baseline_full_rank_evaluation_equals_dense_scores():
    from hassaku_b200.algorithms.sgd_alg import SGDBaseline
    from hassaku_b200.data.dataset import FullEvalDataset
    from hassaku_b200.data.synthetic import make_interactions
    from hassaku_b200.eval.eval import FullEvaluator, evaluate_recommender_algorithm
    data = make_interactions(300, 200, 6000, seed=0, n_user_groups=2)
    ds = FullEvalDataset.from_interactions(data.val, data.train, 'val', data.user_group, 2)
    torch.manual_seed(3)
    m = SGDBaseline(300, 200)
    with torch.no_grad():
        for p in m.parameters():
            p.copy_(torch.randn_like(p))
    m.to('cuda')

    class L:
        dataset, batch_size = ds, 128

    got = evaluate_recommender_algorithm(m, L, FullEvaluator(True, 2, ds.user_to_user_group), 'cuda')
    # dense restatement of the same sweep (eval.py:243-253 with the baseline's combine, sgd_alg.py:99-102)
    scores = m.user_bias.weight.detach() + m.item_bias.weight.detach().view(1, -1) + m.global_bias.detach()
    scores[torch.from_numpy(data.train.toarray().astype(bool)).cuda()] = -torch.inf
    ev = FullEvaluator(True, 2, ds.user_to_user_group)
    ev.eval_batch(torch.arange(300, device='cuda'), scores, torch.from_numpy(data.val.toarray()).float().cuda())
    want = ev.get_results()
    assert sorted(got) == sorted(want)
    for k, v in want.items():
        assert abs(got[k] - v) <= 1e-6, (k, got[k], v)


@pytest.mark.parametrize('aggr', [True, False])
def test_calibration_decorator_over_device_evaluator_vs_reference_fixture(aggr):
    from hassaku_b200.eval.eval import FullEvaluator, FullEvaluatorCalibrationDecorator as Deco
    g = load_golden('calibration_kat')
    ev = FullEvaluator(aggr, 2, torch.from_numpy(g['user_group']))
    ev = Deco(ev, torch.from_numpy(g['item_tag']), torch.from_numpy(g['user_tag']), 'tag', float(g['beta']))
    ev = Deco(ev, torch.from_numpy(g['item_pop']), torch.from_numpy(g['user_pop']), 'pop', float(g['beta']))
    logits, y = torch.from_numpy(g['logits']).cuda(), torch.from_numpy(g['y_true']).cuda()
    for lo, hi in ((0, 24), (24, 40)):
        ev.eval_batch(torch.arange(lo, hi, device='cuda'), logits[lo:hi], y[lo:hi])
    res = ev.get_results()
    if aggr:
        want = dict(zip(g['aggr/names'], g['aggr/values']))
        assert sorted(res) == sorted(want)
        for k, v in want.items():
            assert np.isclose(res[k], v, rtol=1e-5, atol=1e-7, equal_nan=True), (k, res[k], v)
    else:
        for k in g['peruser/names']:
            # atol: the Jensen-Shannon / KL values of near-identical distributions are differences of logs (cancellation):
            # the GPU's logf and the fixture's CPU logf agree to ~1e-6 absolute there (measured 2.2e-6 at a value of 2.7e-3)
            assert np.allclose(res[k], g[f'peruser/{k}'], rtol=1e-5, atol=5e-6, equal_nan=True), k


def test_shard_local_index_matches_formula():
    from hassaku_b200 import _C
    gen = torch.Generator(device='cuda'); gen.manual_seed(0)
    idx = torch.randint(0, 1_000_003, (8192, 51), device='cuda', generator=gen)
    idx[0, 0] = -5                                     # negative indices pass through for the consumer's bounds check
    for world, stride in ((1, 0), (2, 0), (8, 125_011), (3, 7)):
        got = _C.shard_local_index(idx, world, stride)
        want = (idx % world) * stride + torch.div(idx, world, rounding_mode='floor')
        want[0, 0] = -5
        assert torch.equal(got, want)
    out = idx.clone()
    _C.shard_local_index(out, 4, 9, out=out)           # in place
    want = (idx % 4) * 9 + torch.div(idx, 4, rounding_mode='floor'); want[0, 0] = -5
    assert torch.equal(out, want)


def test_factor_model_evaluation_equals_dense_predict_path():
    """A fitted SVD-like model goes through the fused top-k kernels and gives the metrics of the reference's dense
    predict path (eval.py:224-236) on the same factors."""
    from hassaku_b200.data.dataset import FullEvalDataset
    from hassaku_b200.data.synthetic import make_interactions
    from hassaku_b200.eval.eval import FullEvaluator, evaluate_recommender_algorithm
    data = make_interactions(300, 200, 6000, seed=0, n_user_groups=2)
    ds = FullEvalDataset.from_interactions(data.val, data.train, 'val', data.user_group, 2)
    rng = np.random.RandomState(1)

    class SVDLike:
        name = 'SVDAlgorithm'
        users_factors = rng.randn(300, 12)
        items_factors = rng.randn(200, 12)

    class L:
        dataset, batch_size = ds, 128

    alg = SVDLike()
    got = evaluate_recommender_algorithm(alg, L, FullEvaluator(True, 2, ds.user_to_user_group), 'cuda')
    scores = torch.from_numpy((alg.users_factors.astype(np.float32) @ alg.items_factors.astype(np.float32).T)).cuda()
    scores[torch.from_numpy(data.train.toarray().astype(bool)).cuda()] = -torch.inf
    ev = FullEvaluator(True, 2, ds.user_to_user_group)
    ev.eval_batch(torch.arange(300, device='cuda'), scores, torch.from_numpy(data.val.toarray()).float().cuda())
    want = ev.get_results()
    assert sorted(got) == sorted(want)
    for k, v in want.items():
        assert abs(got[k] - v) <= 1e-6, (k, got[k], v)


def test_hit_at_k_through_both_evaluator_paths():
    from hassaku_b200.eval.eval import FullEvaluator
    from hassaku_b200.eval.metrics import hit_at_k_batch
    g = load_golden('calibration_kat')
    logits, y = torch.from_numpy(g['logits']).cuda(), torch.from_numpy(g['y_true']).cuda()
    grp = torch.from_numpy(g['user_group'])
    want = {}
    top = logits.topk(100).indices
    for k in (5, 10, 50, 100):
        hit = (torch.gather(y, 1, top[:, :k]).sum(1) > 0).double().cpu()
        want[f'hit@{k}'] = float(hit.mean())
        for gi in range(2):
            want[f'group_{gi}_hit@{k}'] = float(hit[grp == gi].mean())
        assert abs(float(hit_at_k_batch(logits, y, k)) - float(hit.sum())) < 1e-6
    ev = FullEvaluator(True, 2, grp, hit=True)
    for lo, hi in ((0, 24), (24, 40)):
        ev.eval_batch(torch.arange(lo, hi, device='cuda'), logits[lo:hi], y[lo:hi])
    res = ev.get_results()
    assert len(res) == 48                      # (3 + 1) metrics x 4 k x (ALL + 2 groups)
    for k_, v in want.items():
        assert abs(res[k_] - v) < 1e-9, (k_, res[k_], v)
    ref = dict(zip(g['aggr/names'], g['aggr/values']))
    for k_, v in res.items():
        if 'hit@' not in k_:
            assert abs(v - ref[k_]) < 1e-6, k_   # the reference's own metrics are unchanged by hit=True


@pytest.mark.parametrize('U,I,d,B,N,kind,biases', [
    (6040, 3706, 402, 512, 50, 'bpr', (False, True, False)),     # cfg2 shape: NV 4 with the scalar tail round
    (6040, 3706, 402, 128, 50, 'bpr', (False, True, False)),     # small batch: item slots split over gridDim.y
    (1000, 500, 256, 64, 9, 'bce', (True, True, True)),
    (300, 200, 7, 33, 3, 'bpr', (True, False, True)),
    (300, 200, 1024, 17, 5, 'bce', (False, False, False)),
    (5000, 3000, 128, 256, 50, 'bpr', (False, True, False)),
])
def test_ring_kernel_equals_the_register_gather_kernel(U, I, d, B, N, kind, biases, monkeypatch):
    """HSK_TRAIN_RING (hsk_train_tma.cu) against HSK_TRAIN_REGS (hsk_train.cu) on the same inputs: scores, dL/ds, loss and
    every gradient table."""
    from hassaku_b200 import _C
    from hassaku_b200.algorithms.sgd_alg import ArenaLayout
    lay = ArenaLayout(U, I, d, *biases)
    gen = torch.Generator(device='cuda'); gen.manual_seed(0)
    arena = torch.zeros(lay.n_total, device='cuda')
    for v, scale in zip(lay.views(arena), (d ** -0.5, d ** -0.5, 0.1, 0.1, 0.1)):
        if v is not None:
            v.copy_(torch.randn(v.shape, device='cuda', generator=gen) * scale)
    u = torch.randint(0, U, (B,), device='cuda', generator=gen)
    i = torch.randint(0, I, (B, N + 1), device='cuda', generator=gen)
    i[:, 1] = i[:, 2]
    u[:4] = u[0]
    out = {}
    for variant in ('regs', 'ring'):
        monkeypatch.setattr(_C, 'TRAIN_VARIANT', variant)
        g = torch.zeros_like(arena); loss = torch.zeros(1, dtype=torch.float64, device='cuda')
        sc = torch.empty((B, N + 1), device='cuda'); ds = torch.empty((B, N + 1), device='cuda')
        st = torch.zeros(1, dtype=torch.int32, device='cuda')
        _C.mf_train_fused(lay.tables(arena), lay.tables(g), u, i, _C.LOSS_KINDS[kind], 0.0, loss, scores_out=sc,
                          dscores_out=ds, status=st)
        assert int(st.item()) == 0
        out[variant] = (sc, ds, loss.item(), g)
    a, b = out['regs'], out['ring']
    assert rel_err(b[0].cpu().numpy(), a[0].cpu().numpy()) < 1e-6
    assert rel_err(b[1].cpu().numpy(), a[1].cpu().numpy()) < 1e-5                # rcp.approx vs rcp.rn, summation order
    assert abs(a[2] - b[2]) <= 1e-6 * abs(a[2])
    assert rel_err(b[3].cpu().numpy(), a[3].cpu().numpy()) < 1e-5
    # an out-of-range item index is reported, not dereferenced
    i_bad = i.clone(); i_bad[3, 2] = I + 5
    monkeypatch.setattr(_C, 'TRAIN_VARIANT', 'ring')
    st = torch.zeros(1, dtype=torch.int32, device='cuda')
    _C.mf_train_fused(lay.tables(arena), lay.tables(torch.zeros_like(arena)), u, i_bad, _C.LOSS_KINDS[kind], 0.0,
                      torch.zeros(1, dtype=torch.float64, device='cuda'), status=st)
    assert int(st.item()) & _C.STATUS_BAD_INDEX


@pytest.mark.parametrize('T', [3, 20, 70])
def test_topk_tag_means_matches_the_reference_gather(T):
    """hsk_topk_tag_means = item_tag_mtx[top ids][:, :k].sum(1) / k of eval/eval.py:174-179 for every k at once (fp64 running
    sums, one rounding), padding ids (-1) skipped."""
    from hassaku_b200 import _C
    torch.manual_seed(T)
    B, I, kl = 257, 5000, 100
    tag = torch.rand(I, T, device='cuda')
    tag = tag / tag.sum(-1, keepdim=True)
    ids = torch.stack([torch.randperm(I, device='cuda')[:kl] for _ in range(B)]).to(torch.int32)
    ks = [100, 50, 10, 5]
    got = _C.topk_tag_means(ids, tag, ks)
    rows = tag.double()[ids.long()]
    for t, k in enumerate(ks):
        want = rows[:, :k].sum(1) / k
        assert float((got[:, t].double() - want).abs().max()) < 1e-7
    ids[3, 40:] = -1                                  # a short list: the divisor stays k (the reference has no short lists)
    got = _C.topk_tag_means(ids, tag, ks)
    want = rows[3, :40].sum(0) / 100
    assert float((got[3, 0].double() - want).abs().max()) < 1e-7
    st = torch.zeros(1, dtype=torch.int32, device='cuda')
    ids[5, 0] = I + 7
    _C.topk_tag_means(ids, tag, ks, status=st)
    assert int(st.item()) & _C.STATUS_BAD_INDEX
