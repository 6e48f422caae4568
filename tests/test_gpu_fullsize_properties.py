"""GPU, BASELINE.json's FULL sizes (where the CPU oracle would take minutes to hours): size-independent properties of the
kernels — two independent kernel paths agree, linearity of the gradient in dL/ds, sortedness / idempotence / shard- and
split-invariance of the top-k, exclusions never surface, dense AdamW touches every element, lazy AdamW only touched rows."""
import math

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _rel(a, b):
    return float((a.double() - b.double()).abs().max() / b.double().abs().max().clamp_min(1e-30))


def _model(U, I, d, seed=0, **kw):
    from hassaku_b200.algorithms.sgd_alg import ArenaLayout

    class M:   # arena-only model built directly on the device (a 2 M x 128 nn.Embedding init on the host is slow)
        pass
    m = M()
    m.layout = ArenaLayout(U, I, d, False, True, False)
    g = torch.Generator(device='cuda'); g.manual_seed(seed)
    m.arena = torch.zeros(m.layout.n_total, device='cuda')
    Uw, Vw, _, Ib, _ = m.layout.views(m.arena)
    Uw.copy_(torch.randn(Uw.shape, device='cuda', generator=g) / math.sqrt(d))
    Vw.copy_(torch.randn(Vw.shape, device='cuda', generator=g) / math.sqrt(d))
    Ib.copy_(torch.randn(Ib.shape, device='cuda', generator=g) * 0.1)
    m.parameters = lambda: [torch.nn.Parameter(m.arena[:4])]
    return m


@pytest.mark.parametrize('U,I,d,B,N,kind', [
    (6040, 3706, 402, 8192, 50, 'bpr'),                  # cfg2, full batch
    (69878, 10677, 128, 8192, 100, 'sampled_softmax'),   # cfg3, full batch
    (2_000_000, 1_000_000, 128, 8192, 50, 'bpr'),        # cfg4 tables on one GPU
])
def test_fused_step_equals_unfused_kernel_path_and_is_linear(U, I, d, B, N, kind):
    """hsk_mf_train_fused (one pass) against hsk_mf_scores -> hsk_rec_loss -> hsk_mf_scatter_grads (three other kernels)."""
    from hassaku_b200 import _C
    m = _model(U, I, d)
    lay = m.layout
    tabs = lay.tables(m.arena)
    gen = torch.Generator(device='cuda'); gen.manual_seed(1)
    u = torch.randint(0, U, (B,), device='cuda', generator=gen)
    i = torch.randint(0, I, (B, N + 1), device='cuda', generator=gen)
    i[:, 1] = i[:, 2]                       # duplicates inside rows
    u[:64] = u[0]                           # a hot user
    kid = _C.LOSS_KINDS[kind]
    shift = math.log(I / N) if kind == 'sampled_softmax' else 0.0
    g1 = torch.zeros_like(m.arena); l1 = torch.zeros(1, dtype=torch.float64, device='cuda')
    sc1 = torch.empty((B, N + 1), device='cuda'); ds1 = torch.empty((B, N + 1), device='cuda')
    _C.mf_train_fused(tabs, lay.tables(g1), u, i, kid, shift, l1, scores_out=sc1, dscores_out=ds1)
    sc2 = torch.empty((B, N + 1), device='cuda'); ds2 = torch.empty((B, N + 1), device='cuda')
    g2 = torch.zeros_like(m.arena); l2 = torch.zeros(1, dtype=torch.float64, device='cuda')
    _C.mf_scores(tabs, u, i, sc2)
    _C.rec_loss(sc2, None, kid, shift, 1.0, l2, ds2)
    _C.mf_scatter_grads(tabs, lay.tables(g2), u, i, ds2)
    assert _rel(sc1, sc2) < 1e-6 and _rel(ds1, ds2) < 1e-5
    assert abs(l1.item() - l2.item()) <= 1e-6 * abs(l2.item())
    assert _rel(g1, g2) < 1e-5
    # linearity of the scatter in dL/ds: scatter(2.5 ds) == 2.5 scatter(ds)
    g3 = torch.zeros_like(m.arena)
    _C.mf_scatter_grads(tabs, lay.tables(g3), u, i, (ds2 * 2.5).contiguous())
    assert _rel(g3, g2 * 2.5) < 1e-5
    # rows no sample touches have exactly-zero gradient
    gU, gV, _, gIb, _ = lay.views(g1)
    untouched = torch.ones(I, dtype=torch.bool, device='cuda'); untouched[i.flatten()] = False
    if bool(untouched.any()):
        assert float(gV[untouched].abs().max()) == 0.0 and float(gIb.view(-1)[untouched].abs().max()) == 0.0
    # bpr: the gradients of a sample's slots sum to zero (the loss only sees score differences)
    if kind == 'bpr':
        assert float(ds1.sum(1).abs().max()) < 1e-9


def test_dense_vs_lazy_adamw_coverage_at_cfg4_size():
    from hassaku_b200 import _C
    from hassaku_b200.train.optim import DenseAdam
    U, I, d, B, N = 2_000_000, 1_000_000, 128, 8192, 50
    gen = torch.Generator(device='cuda'); gen.manual_seed(2)
    u = torch.randint(0, U, (B,), device='cuda', generator=gen)
    i = torch.randint(0, I, (B, N + 1), device='cuda', generator=gen)
    acc = torch.zeros(1, dtype=torch.float64, device='cuda')
    for mode in ('dense', 'lazy'):
        m = _model(U, I, d)
        before = m.arena.clone()
        opt = DenseAdam(m, lr=1e-3, weight_decay=1e-2, mode=mode)
        _C.mf_train_fused(m.layout.tables(m.arena), opt.grad_tables, u, i, 0, 0.0, acc)
        if mode == 'lazy':
            opt.mark(u, i)
        opt.step_fused()
        changed = (m.arena != before)
        Uw_c, Vw_c, _, Ib_c, _ = m.layout.views(changed)
        touched_i = torch.zeros(I, dtype=torch.bool, device='cuda'); touched_i[i.flatten()] = True
        touched_u = torch.zeros(U, dtype=torch.bool, device='cuda'); touched_u[u] = True
        if mode == 'dense':     # decoupled decay moves EVERY non-zero element every step
            z = m.layout.views(before == 0)      # randn does return a few exact zeros among 3.8e8 draws
            assert bool((Vw_c | z[1])[~touched_i].all()) and bool((Uw_c | z[0])[~touched_u].all())
            # touched rows: decay and a tiny-gradient Adam step can cancel exactly in a handful of elements
            assert float(Vw_c[touched_i].float().mean()) > 0.9999 and float(Uw_c[touched_u].float().mean()) > 0.9999
        else:                   # lazy: exactly the touched rows move
            assert bool(Vw_c[touched_i].any(1).all()) and not bool(Vw_c[~touched_i].any())
            assert bool(Uw_c[touched_u].any(1).all()) and not bool(Uw_c[~touched_u].any())
        assert float(opt.g.abs().max()) == 0.0
        del m, opt, before, changed


@pytest.mark.parametrize('prec', ['fp32', 'bf16'])
def test_topk_properties_at_cfg4_item_count(prec):
    """8 192 users x 1 M items, d 128: sorted output, idempotent, exclusions never surface, item-shard invariance."""
    from scipy import sparse as sp
    from hassaku_b200 import _C
    from hassaku_b200.eval.eval import DeviceCSR
    U, I, d, B, k = 8192, 1_000_000, 128, (8192 if prec == 'bf16' else 2048), 100
    m = _model(U, I, d, seed=3)
    lay = m.layout
    Uw, Vw, _, Ib, _ = lay.views(m.arena)
    rng = np.random.RandomState(0)
    rows = np.repeat(np.arange(U), 80)
    ex = sp.csr_matrix((np.ones(len(rows), dtype=bool), (rows, rng.randint(0, I, len(rows)))), shape=(U, I))
    ex.sum_duplicates(); ex.sort_indices()
    exd = DeviceCSR(ex, 'cuda')
    users = torch.arange(B, device='cuda')
    P = _C.PRECISIONS[prec]

    def run(Vw_, Ib_, id_offset=0, id_stride=1):
        s = torch.empty((B, k), device='cuda'); ids = torch.empty((B, k), dtype=torch.int32, device='cuda')
        if prec == 'fp32':
            t = _C.make_tables(Uw, Vw_, None, Ib_, None, d)
            scr = torch.empty(_C.eval_topk_scratch_bytes(B, Vw_.shape[0], k), dtype=torch.uint8, device='cuda')
            _C.eval_topk(t, users, k, s, ids, scr, exd.indptr, exd.indices, id_offset=id_offset, id_stride=id_stride)
        else:
            Uq = _C.pack_rows(Uw, d, P, row_idx=users)
            Vq = _C.pack_rows(Vw_, d, P)
            scr = torch.empty(_C.eval_topk_tc_scratch_bytes(B, Vw_.shape[0], k), dtype=torch.uint8, device='cuda')
            _C.eval_topk_tc(Uq, Vq, P, users, U, k, s, ids, scr, Ib=Ib_, excl_indptr=exd.indptr, excl_indices=exd.indices,
                            id_offset=id_offset, id_stride=id_stride)
        return s, ids

    s1, i1 = run(Vw, Ib)
    s2, i2 = run(Vw, Ib)
    assert torch.equal(i1, i2) and torch.equal(s1, s2)                       # idempotent / deterministic
    assert bool((s1[:, 1:] <= s1[:, :-1]).all())                             # sorted by score
    tie = s1[:, 1:] == s1[:, :-1]
    assert bool((i1[:, 1:][tie] > i1[:, :-1][tie]).all())                    # ties by ascending item id
    assert int((i1 < 0).sum()) == 0 and int(i1.max()) < I
    srt = torch.sort(i1.long(), dim=1).values
    assert bool((srt[:, 1:] != srt[:, :-1]).all())                           # no item twice
    keys = set((ex.tocoo().row.astype(np.int64) * I + ex.tocoo().col).tolist())
    got = (users.cpu().numpy()[:256, None] * I + i1.cpu().numpy()[:256].astype(np.int64)).ravel()
    assert not any(int(x) in keys for x in got)                              # excluded items never surface
    # the returned scores are the model's scores of the returned items
    ref = (Uw[users][:, None, :] * Vw[i1.long()[:, :8]]).sum(-1) + Ib.view(-1)[i1.long()[:, :8]]
    tol = 1e-5 if prec == 'fp32' else 2 ** -7
    assert float((s1[:, :8] - ref).abs().max()) <= tol * float(ref.abs().max())
    # item-shard invariance: score two (i mod 2) shards, merge -> same ids (same order-independent key ordering)
    parts = [run(Vw[r::2].contiguous(), Ib.view(-1)[r::2].contiguous().view(-1, 1), id_offset=r, id_stride=2) for r in range(2)]
    ms = torch.empty((B, k), device='cuda'); mi = torch.empty((B, k), dtype=torch.int32, device='cuda')
    _C.topk_merge(torch.stack([p[0] for p in parts]), torch.stack([p[1] for p in parts]), ms, mi)
    assert torch.equal(mi, i1) and torch.equal(ms, s1)
