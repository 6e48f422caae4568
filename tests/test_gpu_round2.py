"""GPU (1 device): the kernels added in round 2 against torch / oracle restatements —
  hsk_adamw_dense_rows + hsk_mark_batch   bitwise equal to hsk_adamw_dense (hence to torch.optim.AdamW on CUDA)
  hsk_rescore_topk                        fp32 re-scoring of tensor-core candidates vs a CPU fp32 matmul
  hsk_route_items / hsk_shard_pack / hsk_shard_unpack_add   vs their torch contracts (tests/test_sharded_gloo.py)
  ShardedMF at world 1                    the whole sparse / dense / graph-captured step vs the single-GPU step
  cfg5-shaped tensor-core evaluation      64 users x 1 M items x 256 against a CPU fp32 oracle
"""
import math

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def rel_err(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))


# ---------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize('U,I,d,B,N,flags', [
    (5000, 3000, 128, 64, 5, (False, True, False)),      # sparse batch: both tables stamped
    (5000, 3000, 402, 64, 5, (True, True, True)),        # ld 404 (101 float4 per row): rows straddle thread blocks
    (300, 200, 7, 512, 50, (False, True, False)),        # batch covers the tables: plain dense pass is chosen
])
def test_row_stamped_adamw_is_bitwise_the_dense_adamw(U, I, d, B, N, flags):
    """Two copies of one model take the same 6 steps on the SAME gradient (the fused kernel's vector reductions are not
    bitwise reproducible run to run, so the gradient arena of the first copy is copied to the second); one optimizer skips
    the gradient traffic of untouched rows (hsk_adamw_dense_rows), the other streams everything (hsk_adamw_dense): p, m,
    v must be bit-identical and the gradient arena all-zero after every step."""
    from hassaku_b200 import _C
    from hassaku_b200.algorithms.sgd_alg import SGDMatrixFactorization
    from hassaku_b200.train.optim import DenseAdam
    torch.manual_seed(0)
    a = SGDMatrixFactorization(U, I, d, *flags)
    b = SGDMatrixFactorization(U, I, d, *flags)
    b.load_state_dict(a.state_dict())
    a.to('cuda'); b.to('cuda')
    oa = DenseAdam(a, lr=1e-2, weight_decay=1e-3)
    ob = DenseAdam(b, lr=1e-2, weight_decay=1e-3)
    gen = torch.Generator(device='cuda'); gen.manual_seed(1)
    acc = torch.zeros(1, dtype=torch.float64, device='cuda')
    used_rows_kernel = False
    for s in range(6):
        u = torch.randint(0, U, (B,), device='cuda', generator=gen)
        i = torch.randint(0, I, (B, N + 1), device='cuda', generator=gen)
        _C.mf_train_fused(a._tables(), oa.grad_tables, u, i, 0, 0.0, acc)
        ob.g.copy_(oa.g)
        oa.mark(u, i)
        used_rows_kernel |= len(oa._segments) > 0
        oa.step_fused()
        ob.step_fused()
        assert torch.equal(a.arena, b.arena) and torch.equal(oa.m, ob.m) and torch.equal(oa.v, ob.v), s
        assert float(oa.g.abs().max()) == 0.0 and float(ob.g.abs().max()) == 0.0
    assert used_rows_kernel == (B * (N + 1) <= I or 4 * B <= U)
    # stamps wrap after 255 steps without losing gradients: jump the step counter across the wrap
    if used_rows_kernel:
        for t0 in (254, 509):
            oa.t = ob.t = t0
            for s in range(3):
                u = torch.randint(0, U, (B,), device='cuda', generator=gen)
                i = torch.randint(0, I, (B, N + 1), device='cuda', generator=gen)
                _C.mf_train_fused(a._tables(), oa.grad_tables, u, i, 0, 0.0, acc)
                ob.g.copy_(oa.g)
                oa.mark(u, i)
                oa.step_fused()
                ob.step_fused()
                assert torch.equal(a.arena, b.arena) and float(oa.g.abs().max()) == 0.0


def test_rescore_topk_against_cpu_fp32():
    from hassaku_b200 import _C
    U, I, d, Be, n_cand, k, G, r = 300, 5000, 200, 77, 128, 100, 3, 1
    gen = torch.Generator().manual_seed(0)
    Uw = torch.randn((U, d), generator=gen) / math.sqrt(d)
    Vfull = torch.randn((I, d), generator=gen) / math.sqrt(d)
    Ibfull = torch.randn(I, generator=gen) * 0.1
    Ub = torch.randn(U, generator=gen) * 0.1
    Gb = torch.tensor([0.3])
    Vloc, Ibloc = Vfull[r::G].contiguous(), Ibfull[r::G].contiguous()      # shard r of G: global id = r + local * G
    users = torch.randperm(U, generator=gen)[:Be]
    n_loc = Vloc.shape[0]
    cand = torch.stack([torch.randperm(n_loc, generator=gen)[:n_cand] * G + r for _ in range(Be)]).to(torch.int32)
    cand[3, 100:] = -1                       # short list
    cand[5, :] = -1                          # empty list
    t = _C.make_tables(Uw.cuda(), Vloc.cuda(), Ub.cuda(), Ibloc.cuda(), Gb.cuda(), d)
    keep = (Uw.cuda(), Vloc.cuda(), Ub.cuda(), Ibloc.cuda(), Gb.cuda())
    t = _C.make_tables(keep[0], keep[1], keep[2], keep[3], keep[4], d)
    s = torch.empty((Be, k), device='cuda'); ids = torch.empty((Be, k), dtype=torch.int32, device='cuda')
    st = torch.zeros(1, dtype=torch.int32, device='cuda')
    _C.rescore_topk(t, users.cuda(), cand.cuda(), k, s, ids, id_offset=r, id_stride=G, status=st)
    assert int(st.item()) == 0
    s, ids = s.cpu(), ids.cpu()
    for b in range(Be):
        c = cand[b][cand[b] >= 0].long()
        sc = (Vfull[c].double() @ Uw[users[b]].double()) + Ub[users[b]].double() + Ibfull[c].double() + 0.3
        order = np.lexsort((c.numpy(), -sc.numpy()))[:k]
        n = len(order)
        assert ids[b, :n].tolist() == c[order].tolist(), b
        assert n == 0 or rel_err(s[b, :n].numpy(), sc[order].numpy()) < 1e-5
        assert (ids[b, n:] == -1).all() and torch.isinf(s[b, n:]).all()
    # an id that is not of this shard is reported, not dereferenced
    bad = cand.clone(); bad[0, 0] = r + 1
    _C.rescore_topk(t, users.cuda(), bad.cuda(), k, torch.empty((Be, k), device='cuda'),
                    torch.empty((Be, k), dtype=torch.int32, device='cuda'), id_offset=r, id_stride=G, status=st)
    assert int(st.item()) & _C.STATUS_BAD_INDEX


# ---------------------------------------------------------------------------------------------------------------
def _route_ref(i_global, n_items, G, capq, ld):
    br = capq + math.ceil(capq / ld)
    flat = i_global.reshape(-1).cpu()
    req = torch.full((G, capq), -1, dtype=torch.int32)
    cnt = torch.zeros(G, dtype=torch.int32)
    out = torch.full_like(flat, -1)
    for q in range(G):
        sel = (flat % G) == q
        ids_all = torch.unique(flat[sel], sorted=True)
        ids = ids_all[:capq]
        req[q, :len(ids)] = (ids // G).to(torch.int32)
        cnt[q] = len(ids)
        if len(ids):
            pos = torch.searchsorted(ids, flat[sel])
            ok = (pos < len(ids)) & (ids[pos.clamp(max=len(ids) - 1)] == flat[sel])
            out[sel] = torch.where(ok, q * br + pos, torch.full_like(pos, -1))
    return req, cnt, out.view(i_global.shape)


@pytest.mark.parametrize('n_items,G,B,N1,capq', [
    (1_000_003, 8, 8192, 51, None),       # cfg4-like, ragged shards
    (1_000_003, 1, 4096, 51, None),
    (3706, 2, 8192, 51, None),            # every item requested
    (41, 3, 24, 6, None),
    (100_000, 4, 2048, 51, 5000),         # capacity overflow: flagged, overflowing ids -> -1
])
def test_route_items_matches_the_torch_contract(n_items, G, B, N1, capq):
    from hassaku_b200 import _C
    from hassaku_b200.sharded import exchange_capacity
    ld = 128
    gen = torch.Generator(device='cuda'); gen.manual_seed(0)
    i = torch.randint(0, n_items, (B, N1), device='cuda', generator=gen)
    i[:, 2] = i[:, 1]
    cq = capq or exchange_capacity(B * N1, n_items, G)
    req = torch.zeros((G, cq), dtype=torch.int32, device='cuda')
    cnt = torch.zeros(G, dtype=torch.int32, device='cuda')
    comp = torch.empty_like(i)
    scr = torch.empty(_C.route_scratch_bytes(n_items, G), dtype=torch.uint8, device='cuda')
    st = torch.zeros(1, dtype=torch.int32, device='cuda')
    for _ in range(2):      # twice: the scratch is reused without clearing by the caller
        _C.route_items(i, n_items, G, cq, ld, req, cnt, comp, scr, st)
    w_req, w_cnt, w_comp = _route_ref(i, n_items, G, cq, ld)
    assert torch.equal(req.cpu(), w_req) and torch.equal(cnt.cpu(), w_cnt) and torch.equal(comp.cpu(), w_comp)
    overflow = capq is not None
    assert bool(int(st.item()) & _C.STATUS_CAPACITY) == overflow
    assert bool((comp < 0).any()) == overflow
    # a bad index is reported and maps to -1
    i2 = i.clone(); i2[0, 0] = n_items + 3
    st.zero_()
    _C.route_items(i2, n_items, G, cq, ld, req, cnt, comp, scr, st)
    assert int(st.item()) & _C.STATUS_BAD_INDEX and int(comp[0, 0]) == -1


def test_shard_pack_and_unpack_add_match_the_torch_contract():
    from hassaku_b200 import _C
    G, capq, ld, n_local = 3, 37, 8, 50
    br = _C.shard_block_rows(capq, ld)
    assert br == capq + math.ceil(capq / ld)
    gen = torch.Generator().manual_seed(0)
    V = torch.randn((n_local, ld), generator=gen)
    Ib = torch.randn(n_local, generator=gen)
    rows = torch.full((G, capq), -1, dtype=torch.int32)
    for q in range(G):
        n = [37, 10, 0][q]
        rows[q, :n] = torch.sort(torch.randperm(n_local, generator=gen)[:n]).values.to(torch.int32)
    out = torch.zeros((G, br, ld), device='cuda')
    _C.shard_pack(V.cuda(), Ib.cuda(), rows.cuda(), G, capq, out)
    want = torch.zeros((G, br, ld))
    for q in range(G):
        for k in range(capq):
            r = int(rows[q, k])
            if r >= 0:
                want[q, k] = V[r]
                want[q].view(-1)[capq * ld + k] = Ib[r]
    assert torch.equal(out.cpu(), want)
    # the way back: gradients of the same layout are ADDED (two peers may hold the same row), rows stamped
    inp = torch.randn((G, br, ld), generator=gen)
    gV = torch.randn((n_local, ld), generator=gen)
    gIb = torch.randn(n_local, generator=gen)
    stamps = torch.zeros(n_local, dtype=torch.uint8)
    gV_d, gIb_d, st_d = gV.cuda(), gIb.cuda(), stamps.cuda()
    step = 700
    _C.shard_unpack_add(inp.cuda(), rows.cuda(), G, capq, gV_d, gIb_d, st_d, step=step)
    wV, wIb, wst = gV.clone().double(), gIb.clone().double(), stamps.clone()
    for q in range(G):
        for k in range(capq):
            r = int(rows[q, k])
            if r >= 0:
                wV[r] += inp[q, k].double()
                wIb[r] += inp[q].view(-1)[capq * ld + k].double()
                wst[r] = _C.row_stamp(step)
    assert rel_err(gV_d.cpu().numpy(), wV.numpy()) < 1e-6 and rel_err(gIb_d.cpu().numpy(), wIb.numpy()) < 1e-6
    assert torch.equal(st_d.cpu(), wst) and _C.row_stamp(step) == 1 + step % 255


# ---------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize('exchange', ['sparse', 'sparse_graph', 'dense', 'dense_graph', 'peer', 'peer_graph'])
@pytest.mark.parametrize('kind,flags', [('bpr', (False, True, False)), ('sampled_softmax', (True, True, False)),
                                        ('bce', (True, True, True))])
def test_sharded_step_at_world_1_equals_the_single_gpu_step(exchange, kind, flags):
    """ShardedMF with world 1 (no process group: every exchange is the identity) runs the complete routed pipeline —
    hsk_route_items, pack, compact table, fused kernel with global normalisers, unpack-add, row-stamped AdamW, eager and
    as a captured CUDA graph — and must reproduce FusedMFTrainStep on the same batches.  'peer': the step kernel over the
    peer table (hsk_mf_train_fused_peer; one rank = its own memory through the same owner / row addressing, system-scope
    reductions and kernel-written stamps).  (No global bias under sampled
    softmax: its gradient sum_j (softmax_j - [j = 0]) is mathematically zero, so Adam normalises pure rounding noise.)"""
    from hassaku_b200 import _C
    from hassaku_b200.algorithms.sgd_alg import SGDMatrixFactorization
    from hassaku_b200.sharded import ShardedMF
    from hassaku_b200.train.optim import DenseAdam
    from hassaku_b200.train.rec_losses import RecommenderSystemLossesEnum
    from hassaku_b200.train.trainer_step import FusedMFTrainStep
    U, I, d, B, N = 3001, 20011, 128, 256, 20
    torch.manual_seed(3)
    single = SGDMatrixFactorization(U, I, d, *flags)
    with torch.no_grad():
        for p in single.parameters():
            p.copy_(torch.randn_like(p) * (1 / math.sqrt(d) if p.shape[-1] == d else 0.1))
    sd0 = {k_: v.clone() for k_, v in single.state_dict().items()}
    single.to('cuda')
    lr, wd = 1e-3, 1e-4

    class _DS:
        n_items = I

    loss_fn = RecommenderSystemLossesEnum[kind].value.build_from_conf({'train_neg_strategy': 'uniform', 'neg_train': N}, _DS())
    opt = DenseAdam(single, lr=lr, weight_decay=wd)
    step = FusedMFTrainStep(single, loss_fn, opt)
    smf = ShardedMF(U, I, d, *flags, world=1, rank=0, device='cuda')
    smf.load_full_state_dict(sd0)
    shift = float(loss_fn.neg_shift())
    rng = np.random.RandomState(9)
    # the peer step is the quarter-warp kernel: the single-GPU side runs the same kernel, so that gradients which are
    # mathematically zero (the user bias under sampled softmax) carry the same rounding noise into Adam's normalisation
    variant0 = _C.TRAIN_VARIANT
    try:
        for s in range(5):
            u = torch.from_numpy(rng.randint(0, U, B).astype(np.int64)).cuda()
            i = torch.from_numpy(rng.randint(0, I, (B, N + 1)).astype(np.int64)).cuda()
            i[:, 2] = i[:, 1]
            _C.TRAIN_VARIANT = 'q' if exchange.startswith('peer') else variant0
            step(u, i)
            _C.TRAIN_VARIANT = variant0
            smf.step(u, i, B, kind, shift, lr, wd, exchange=exchange)
            l_single, l_sh = step.pop_loss_sum(), smf.pop_loss()
            assert abs(l_single - l_sh) <= 1e-5 * abs(l_single), (s, l_single, l_sh)
        smf.check_status()
        sd = smf.full_state_dict()
        for n, p in single.state_dict().items():
            err = rel_err(sd[n].numpy(), p.detach().cpu().numpy())
            assert err < 1e-5 + 2e-3 * lr, (n, err)
        assert float(smf.g.abs().max()) == 0.0
    finally:
        _C.TRAIN_VARIANT = variant0
        smf.close()


def test_sharded_evaluate_at_world_1_equals_the_single_gpu_sweep():
    from hassaku_b200.algorithms.sgd_alg import SGDMatrixFactorization
    from hassaku_b200.data.dataset import FullEvalDataset
    from hassaku_b200.data.synthetic import make_interactions
    from hassaku_b200.eval.eval import FullEvaluator, evaluate_recommender_algorithm
    from hassaku_b200.sharded import ShardedMF
    U, I, d = 1201, 4007, 64
    torch.manual_seed(5)
    single = SGDMatrixFactorization(U, I, d, True, True, True)
    with torch.no_grad():
        for p in single.parameters():
            p.copy_(torch.randn_like(p) * (0.4 if p.shape[-1] == d else 0.1))
    sd0 = {k_: v.clone() for k_, v in single.state_dict().items()}
    single.to('cuda')
    data = make_interactions(U, I, 40000, seed=2, n_user_groups=2)
    ds = FullEvalDataset.from_interactions(data.val, data.train, 'val', data.user_group, 2)

    class L:
        dataset, batch_size = ds, 256

    smf = ShardedMF(U, I, d, True, True, True, world=1, rank=0, device='cuda')
    smf.load_full_state_dict(sd0)
    for prec in ('fp32', 'bf16', 'tf32'):
        single.eval_precision = prec
        ref = evaluate_recommender_algorithm(single, L, FullEvaluator(True, 2, ds.user_to_user_group), 'cuda')
        got = smf.evaluate(data.val, data.train, FullEvaluator(True, 2, ds.user_to_user_group), batch_size=200, precision=prec)
        rep = smf.evaluate_replicated(data.val, data.train, FullEvaluator(True, 2, ds.user_to_user_group), batch_size=300,
                                      precision=prec)
        assert sorted(got) == sorted(ref) == sorted(rep)
        for k_, v in ref.items():
            assert abs(got[k_] - v) <= 1e-9, (prec, k_, got[k_], v)
            assert abs(rep[k_] - v) <= 1e-9, (prec, 'replicated', k_, rep[k_], v)
    smf.check_status()


# ---------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize('prec', ['bf16', 'tf32'])
def test_tensor_core_evaluator_at_cfg5_shape_against_a_cpu_fp32_oracle(prec):
    """64 users x 1 M items x d 256 (the cfg5 item table), 80 exclusions per user, top-100: the ids of the tensor-core
    mode (with fp32 re-scoring) against a CPU fp32 matmul + mask + sort — the oracle's A.7 on this slice."""
    from scipy import sparse as sp
    from hassaku_b200.algorithms.sgd_alg import SGDMatrixFactorization
    from hassaku_b200.eval.eval import DeviceCSR, TopKScorer
    U, I, d, k = 64, 1_000_000, 256, 100
    torch.manual_seed(11)
    m = SGDMatrixFactorization(U, I, d, use_item_bias=True)
    with torch.no_grad():
        m.user_embeddings.weight.copy_(torch.randn(U, d) / math.sqrt(d))
        m.item_embeddings.weight.copy_(torch.randn(I, d) / math.sqrt(d) * 2.0)
        m.item_bias.weight.copy_(torch.randn(I, 1) * 0.05)
    Uw, Vw, Ib = m.user_embeddings.weight.detach().clone(), m.item_embeddings.weight.detach().clone(), m.item_bias.weight.detach().clone()
    rng = np.random.RandomState(0)
    rows = np.repeat(np.arange(U), 80)
    ex = sp.csr_matrix((np.ones(len(rows), dtype=bool), (rows, rng.randint(0, I, len(rows)))), shape=(U, I))
    ex.sum_duplicates(); ex.sort_indices()
    # CPU oracle: fp32 scores (eval.py:247-248), -inf on exclusions (:250-251), top-k with (score desc, id asc)
    scores = (Uw @ Vw.t()) + Ib.view(1, -1)
    scores[torch.from_numpy(ex.tocoo().row.astype(np.int64)), torch.from_numpy(ex.tocoo().col.astype(np.int64))] = -torch.inf
    top_s, top_i = torch.topk(scores, k + 1, dim=1)
    m.to('cuda')
    s_tc, i_tc = TopKScorer(m, U, k, prec)(torch.arange(U, device='cuda'), DeviceCSR(ex, 'cuda'))
    m.check_status()
    s_tc, i_tc = s_tc.cpu(), i_tc.cpu().long()
    overlap = np.mean([len(np.intersect1d(a, b)) / k for a, b in zip(top_i[:, :k].numpy(), i_tc.numpy())])
    assert overlap >= 0.999, overlap
    assert rel_err(s_tc.numpy(), top_s[:, :k].numpy()) < 1e-5
    # ids identical wherever neighbouring fp32 scores are separated by more than the summation-order noise
    gap = (top_s[:, :-1] - top_s[:, 1:]) > 1e-5 * float(top_s.abs().max())
    clear = gap[:, :k].clone()
    clear[:, 1:] &= gap[:, :k - 1]
    assert bool((i_tc == top_i[:, :k])[clear].all())


@pytest.mark.parametrize('prec', ['bf16', 'tf32'])
@pytest.mark.parametrize('U,I,d,B,n_excl', [
    (700, 50_011, 96, 700, 40),        # 6 row tiles (3 pairs), ragged last item tile, one split
    (300, 20_000, 256, 129, 0),        # 2 row tiles of which the second holds one user; several splits; no exclusions
    (130, 1_000, 64, 130, 300),        # more exclusions than kept entries: every cut checks the exclusion row
    (5, 129, 16, 5, 3),                # fewer items than one tile
])
def test_pair_kernel_matches_the_single_cta_kernel(prec, U, I, d, B, n_excl):
    """cta_group::2 kernel (item bias pre-loaded into TMEM, 256-wide tiles) against the cta_group::1 kernel on the same packed
    operands: same candidates (raw low-precision ranking, k = 128) up to the rounding of bias + dot vs dot + bias."""
    from scipy import sparse as sp
    from hassaku_b200 import _C
    from hassaku_b200.eval.eval import DeviceCSR
    k = min(128, I)
    gen = torch.Generator(device='cuda'); gen.manual_seed(0)
    Uw = torch.randn((U, d), device='cuda', generator=gen) / d ** 0.5
    Vw = torch.randn((I, d), device='cuda', generator=gen) / d ** 0.5
    Ib = torch.randn((I, 1), device='cuda', generator=gen) * 0.1
    Ub = torch.randn((U, 1), device='cuda', generator=gen) * 0.1
    exd = None
    if n_excl:
        rng = np.random.RandomState(0)
        rows = np.repeat(np.arange(U), n_excl)
        ex = sp.csr_matrix((np.ones(len(rows), dtype=bool), (rows, rng.randint(0, I, len(rows)))), shape=(U, I))
        ex.sum_duplicates(); ex.sort_indices()
        exd = DeviceCSR(ex, 'cuda')
    users = torch.arange(B, device='cuda')
    P = _C.PRECISIONS[prec]
    Uq, Vq = _C.pack_rows(Uw, d, P, row_idx=users), _C.pack_rows(Vw, d, P)
    out = {}
    for variant in ('single', 'pair'):
        s = torch.empty((B, k), device='cuda'); ids = torch.empty((B, k), dtype=torch.int32, device='cuda')
        scr = torch.empty(_C.eval_topk_tc_scratch_bytes(B, I, k), dtype=torch.uint8, device='cuda')
        st = torch.zeros(1, dtype=torch.int32, device='cuda')
        _C.eval_topk_tc(Uq, Vq, P, users, U, k, s, ids, scr, Ub=Ub, Ib=Ib, excl_indptr=exd.indptr if exd else None,
                        excl_indices=exd.indices if exd else None, status=st, variant=variant)
        torch.cuda.synchronize()
        assert int(st.item()) == 0
        out[variant] = (s.cpu(), ids.cpu())
    (s1, i1), (s2, i2) = out['single'], out['pair']
    fin = torch.isfinite(s1)
    assert torch.equal(fin, torch.isfinite(s2))
    scale = float(s1[fin].abs().max())
    assert float((s1[fin] - s2[fin]).abs().max()) <= 1e-6 * scale + 1e-6      # same products, one fp32 add in another order
    same = (i1 == i2)[fin].float().mean()
    assert float(same) >= 0.999, float(same)                                   # order flips only between near-equal scores
    ov = np.mean([len(np.intersect1d(a[f], b[f])) / max(int(f.sum()), 1) for a, b, f in zip(i1.numpy(), i2.numpy(), fin.numpy())])
    assert ov >= 0.9995, ov


# ---------------------------------------------------------------------------------------------------------------
class _Guarded:
    """A tensor carved out of a larger allocation whose margins hold a canary pattern: what compute-sanitizer's memcheck would
    report for WRITES (the tool is closed on this GPU pool, profiles/r02_compute_sanitizer_closed.log) is checked by hand."""
    PAD = 4096   # bytes on each side

    def __init__(self, shape, dtype, fill=None):
        n = int(np.prod(shape)) * torch.empty(0, dtype=dtype).element_size()
        n_al = (n + 255) // 256 * 256
        self.raw = torch.full((self.PAD + n_al + self.PAD,), 0xA5, dtype=torch.uint8, device='cuda')
        self.n = n
        self.t = self.raw[self.PAD:self.PAD + n].view(dtype).view(shape)
        if fill is not None:
            self.t.fill_(fill)

    def intact(self):
        lo = self.raw[:self.PAD]
        hi = self.raw[self.PAD + self.n:]
        return bool((lo == 0xA5).all()) and bool((hi == 0xA5).all())


def test_kernels_do_not_write_outside_their_buffers():
    from scipy import sparse as sp
    from hassaku_b200 import _C
    from hassaku_b200.algorithms.sgd_alg import ArenaLayout
    from hassaku_b200.eval.eval import DeviceCSR
    gen = torch.Generator(device='cuda'); gen.manual_seed(0)
    guards = []

    def G(shape, dtype, fill=None):
        g = _Guarded(shape, dtype, fill)
        guards.append(g)
        return g.t

    # ---- train step kernels (three layouts) + row-stamped AdamW ----
    for d, variant, kind in ((402, 'ring', 0), (128, 'q', 1), (100, 'regs', 2)):
        U, I, B, N = 901, 777, 130, 17
        lay = ArenaLayout(U, I, d, True, True, True)
        arena = G((lay.n_total,), torch.float32, 0.0)
        for v_, sc_ in zip(lay.views(arena), (d ** -0.5, d ** -0.5, 0.1, 0.1, 0.1)):
            v_.copy_(torch.randn(v_.shape, device='cuda', generator=gen) * sc_)
        g, m, v = (G((lay.n_total,), torch.float32, 0.0) for _ in range(3))
        u = torch.randint(0, U, (B,), device='cuda', generator=gen)
        i = torch.randint(0, I, (B, N + 1), device='cuda', generator=gen)
        sc, ds = G((B, N + 1), torch.float32), G((B, N + 1), torch.float32)
        acc = G((1,), torch.float64, 0.0)
        su, si = G((U,), torch.uint8, 0), G((I,), torch.uint8, 0)
        _C.mf_train_fused(lay.tables(arena), lay.tables(g), u, i, kind, 0.3, acc, scores_out=sc, dscores_out=ds, variant=variant)
        _C.mark_batch(u, i, U, I, su, si, step=1)
        _C.adamw_dense_rows(arena, m, v, g, [(lay.off_U, U, lay.ld, su), (lay.off_V, I, lay.ld, si)], 1e-3, 0.9, 0.999, 1e-8, 1e-4, 1)
        assert float(g.abs().max()) == 0.0
    # ---- routing kernels ----
    n_items, Gw, capq, ld = 5003, 3, 700, 128
    i = torch.randint(0, n_items, (64, 21), device='cuda', generator=gen)
    req, cnt, comp = G((Gw, capq), torch.int32), G((Gw,), torch.int32), G((64, 21), torch.int64)
    scr = G((_C.route_scratch_bytes(n_items, Gw),), torch.uint8)
    _C.route_items(i, n_items, Gw, capq, ld, req, cnt, comp, scr)
    br = _C.shard_block_rows(capq, ld)
    V, Ib = torch.randn((1700, ld), device='cuda', generator=gen), torch.randn(1700, device='cuda', generator=gen)
    rows = torch.where(req >= 0, req % 1700, req)
    out = G((Gw, br, ld), torch.float32, 0.0)
    _C.shard_pack(V, Ib, rows, Gw, capq, out)
    gV, gIb, st = G((1700, ld), torch.float32, 0.0), G((1700,), torch.float32, 0.0), G((1700,), torch.uint8, 0)
    _C.shard_unpack_add(out, rows, Gw, capq, gV, gIb, st, step=3)
    # ---- evaluator: both tensor-core kernels, several splits and a ragged last tile, then re-scoring ----
    U, I, d, B, k = 300, 40_011, 96, 257, 128
    Uw = torch.randn((U, d), device='cuda', generator=gen) / d ** 0.5
    Vw = torch.randn((I, d), device='cuda', generator=gen) / d ** 0.5
    Ibv = torch.randn((I, 1), device='cuda', generator=gen) * 0.1
    rng = np.random.RandomState(0)
    rws = np.repeat(np.arange(U), 25)
    ex = sp.csr_matrix((np.ones(len(rws), dtype=bool), (rws, rng.randint(0, I, len(rws)))), shape=(U, I))
    ex.sum_duplicates(); ex.sort_indices()
    exd = DeviceCSR(ex, 'cuda')
    users = torch.arange(B, device='cuda')
    for prec in ('bf16', 'tf32'):
        P = _C.PRECISIONS[prec]
        Uq, Vq = _C.pack_rows(Uw, d, P, row_idx=users), _C.pack_rows(Vw, d, P)
        for variant in ('pair', 'single'):
            s, ids = G((B, k), torch.float32), G((B, k), torch.int32)
            scr = G((_C.eval_topk_tc_scratch_bytes(B, I, k),), torch.uint8)
            _C.eval_topk_tc(Uq, Vq, P, users, U, k, s, ids, scr, Ib=Ibv, excl_indptr=exd.indptr, excl_indices=exd.indices,
                            variant=variant)
            t = _C.make_tables(Uw, Vw, None, Ibv, None, d)
            s2, i2 = G((B, 100), torch.float32), G((B, 100), torch.int32)
            _C.rescore_topk(t, users, ids, 100, s2, i2, cand_scores=s)
    s, ids = G((B, 100), torch.float32), G((B, 100), torch.int32)
    scr = G((_C.eval_topk_scratch_bytes(B, I, 100),), torch.uint8)
    _C.eval_topk(_C.make_tables(Uw, Vw, None, Ibv, None, d), users, 100, s, ids, scr, exd.indptr, exd.indices)
    torch.cuda.synchronize()
    bad = [j for j, g_ in enumerate(guards) if not g_.intact()]
    assert not bad, f'canaries overwritten around buffers {bad}'


@pytest.mark.parametrize('S,I,Be', [(3, 10007, 200), (8, 70001, 64), (2, 513, 300)])
def test_evaluator_over_item_shards_equals_the_single_table(S, I, Be):
    """hsk_eval_topk_tc_shards + hsk_rescore_topk_shards over S interleaved item shards (item i = row i / S of shard i % S;
    separate allocations here, peer-mapped tables in ShardedMF.evaluate_streamed) return exactly what the single-table calls
    return: same candidate ids and low-precision scores, same fp32 top-k — ragged shards, exclusions, item bias, tile
    boundaries inside and between shards, several splits (small batch) included."""
    from hassaku_b200 import _C
    torch.manual_seed(S * 1000 + Be)
    d, k, kc = 64, 20, 48
    P = _C.PRECISIONS['bf16']
    U = 400
    Uw = (torch.randn(U, d, device='cuda') * 0.4).contiguous()
    Vw = (torch.randn(I, d, device='cuda') * 0.4).contiguous()
    Ib = (torch.randn(I, device='cuda') * 0.2).contiguous()
    Ub = (torch.randn(U, device='cuda') * 0.1).contiguous()
    users = torch.randperm(U, device='cuda')[:Be].contiguous()
    # exclusion CSR: ~30 sorted random items per user
    rng = np.random.RandomState(S)
    cols = [np.unique(rng.randint(0, I, 30)) for _ in range(U)]
    indptr = torch.from_numpy(np.concatenate([[0], np.cumsum([len(c) for c in cols])]).astype(np.int64)).cuda()
    indices = torch.from_numpy(np.concatenate(cols).astype(np.int32)).cuda()
    Uq = _C.pack_rows(Uw, d, P, row_idx=users)
    # single table
    Vq = _C.pack_rows(Vw, d, P)
    s1, i1 = torch.empty((Be, kc), device='cuda'), torch.empty((Be, kc), dtype=torch.int32, device='cuda')
    scr = torch.empty(_C.eval_topk_tc_scratch_bytes(Be, I, kc), dtype=torch.uint8, device='cuda')
    _C.eval_topk_tc(Uq, Vq, P, users, U, kc, s1, i1, scr, Ub=Ub, Ib=Ib, excl_indptr=indptr, excl_indices=indices, variant='pair')
    t = _C.make_tables(Uw, Vw, Ub.view(-1, 1), Ib.view(-1, 1), None, d)
    r1s, r1i = torch.empty((Be, k), device='cuda'), torch.empty((Be, k), dtype=torch.int32, device='cuda')
    _C.rescore_topk(t, users, i1, k, r1s, r1i, cand_scores=s1)
    # S shards
    Vs = [Vw[q::S].contiguous() for q in range(S)]
    Ibs = [Ib[q::S].contiguous() for q in range(S)]
    Vqs = [_C.pack_rows(v, d, P) for v in Vs]
    shards = _C.ItemShards([v.shape[0] for v in Vs], Vq=[x.data_ptr() for x in Vqs], V=[v.data_ptr() for v in Vs],
                           Ib=[b.data_ptr() for b in Ibs])
    s2, i2 = torch.empty((Be, kc), device='cuda'), torch.empty((Be, kc), dtype=torch.int32, device='cuda')
    scr2 = torch.empty(_C.eval_topk_tc_shards_scratch_bytes(Be, shards, kc), dtype=torch.uint8, device='cuda')
    st = torch.zeros(1, dtype=torch.int32, device='cuda')
    _C.eval_topk_tc_shards(Uq, shards, P, users, U, kc, s2, i2, scr2, Ub=Ub, excl_indptr=indptr, excl_indices=indices, status=st)
    assert int(st.item()) == 0
    assert torch.equal(i1, i2), 'candidate ids differ'
    assert torch.equal(s1, s2), 'candidate scores differ'
    r2s, r2i = torch.empty((Be, k), device='cuda'), torch.empty((Be, k), dtype=torch.int32, device='cuda')
    _C.rescore_topk_shards(t, shards, users, i2, k, r2s, r2i, status=st, cand_scores=s2)
    assert int(st.item()) == 0
    assert torch.equal(r1i, r2i) and torch.equal(r1s, r2s)
    # none of the returned ids is excluded
    ex = {(u, int(c)) for u in range(U) for c in cols[u]}
    got = r2i.cpu().numpy()
    for b, u in enumerate(users.cpu().numpy()[:50]):
        assert not any((int(u), int(x)) in ex for x in got[b])
