"""CPU, world_size 2, gloo: the routing of the item-/user-sharded training step (hassaku_b200/sharded.py) — dedupe,
owner grouping, the three all-to-all exchanges, global normalisers, replicated global bias — reproduces the
single-process oracle step on the union batch.  The arithmetic is plugged in as a torch reference here (the CUDA
kernels need a GPU; their own parity is tests/test_gpu_*.py), so this isolates the distributed host logic."""
import math
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from hsk_testutil import ROOT  # noqa: F401


class TorchRefOps:
    """Test-only stand-in for hassaku_b200.sharded.CudaOps: torch restatements of the routing kernels' contracts
    (include/hassaku_b200.h: hsk_route_items, hsk_shard_pack, hsk_shard_unpack_add, hsk_mark_rows) and of the arithmetic."""

    def local_index(self, idx, world, rank_stride):
        return (idx % world) * rank_stride + torch.div(idx, world, rounding_mode='floor')

    def route(self, i_global, n_items, world, capq, ld, req_rows, req_count, compact_idx, scratch):
        br = capq + math.ceil(capq / ld)
        flat = i_global.reshape(-1)
        out = torch.full_like(flat, -1)
        req_rows.fill_(-1)
        for q in range(world):
            sel = (flat % world) == q
            ids = torch.unique(flat[sel], sorted=True)[:capq]          # ascending local rows
            req_rows[q, :len(ids)] = torch.div(ids, world, rounding_mode='floor').to(torch.int32)
            req_count[q] = len(ids)
            pos = torch.searchsorted(ids, flat[sel])
            ok = (pos < len(ids)) & (ids[pos.clamp(max=max(len(ids) - 1, 0))] == flat[sel]) if len(ids) else torch.zeros_like(pos, dtype=torch.bool)
            out[sel] = torch.where(ok, q * br + pos, torch.full_like(pos, -1))
        compact_idx.copy_(out.view_as(compact_idx))

    def pack(self, V2d, Ib, rows, world, capq, out):
        ld = V2d.shape[1]
        for q in range(world):
            for k in range(capq):
                r = int(rows[q, k])
                if r >= 0:
                    out[q, k] = V2d[r]
                    if Ib is not None:
                        out[q].view(-1)[capq * ld + k] = Ib[r]

    def unpack_add(self, inp, rows, world, capq, gV2d, gIb, stamps, step, step_dev):
        ld = gV2d.shape[1]
        for q in range(world):
            for k in range(capq):
                r = int(rows[q, k])
                if r >= 0:
                    gV2d[r] += inp[q, k]
                    if gIb is not None:
                        gIb[r] += inp[q].view(-1)[capq * ld + k]
                    stamps[r] = 1 + step % 255

    def mark_rows(self, idx, n_rows, stamps, step, step_dev):
        stamps[idx] = 1 + step % 255

    def train_fused(self, lay, arena, g_arena, Vc, Ibc, gVc, gIbc, u_local, compact_idx, B_global, kind, shift, loss_accum):
        Uw, _, Ub, _, Gb = lay.views(arena)
        gU, _, gUb, _, gGb = lay.views(g_arena)
        leaves = {'U': Uw.detach().clone().requires_grad_(), 'V': Vc[:, :lay.d].detach().clone().requires_grad_()}
        if Ub is not None:
            leaves['Ub'] = Ub.detach().clone().requires_grad_()
        if Ibc is not None:
            leaves['Ib'] = Ibc.detach().clone().requires_grad_()
        if Gb is not None:
            leaves['Gb'] = Gb.detach().clone().requires_grad_()
        s = (leaves['U'][u_local][:, None, :] * leaves['V'][compact_idx]).sum(-1)
        if 'Ub' in leaves:
            s = s + leaves['Ub'][u_local]
        if 'Ib' in leaves:
            s = s + leaves['Ib'][compact_idx]
        if 'Gb' in leaves:
            s = s + leaves['Gb']
        B, N1 = compact_idx.shape
        if kind == 0:
            x = s[:, :1] - s[:, 1:]
            loss = (-torch.nn.functional.logsigmoid(x)).double().sum() / (B_global * (N1 - 1))
        elif kind == 1:
            s2 = torch.cat([s[:, :1], s[:, 1:] + shift], 1)
            loss = (-s2[:, 0] + torch.logsumexp(s2, -1)).double().sum() / B_global
        else:
            y = torch.zeros_like(s); y[:, 0] = 1.
            loss = torch.nn.functional.binary_cross_entropy_with_logits(s, y, reduction='sum').double() / (B_global * N1)
        loss.backward()
        gU += leaves['U'].grad
        gVc[:, :lay.d] += leaves['V'].grad
        if 'Ub' in leaves:
            gUb += leaves['Ub'].grad
        if 'Ib' in leaves:
            gIbc += leaves['Ib'].grad
        if 'Gb' in leaves:
            gGb += leaves['Gb'].grad
        loss_accum += loss.detach()

    def adamw(self, arena, m, v, g, segments, lr, wd, t, decoupled=True, consts_dev=None, step_dev=None):
        # contract of hsk_adamw_dense_rows: the gradient is zero outside the rows stamped this step
        for off, rows, ld, stamps in segments:
            untouched = stamps[:rows] != 1 + t % 255
            assert float(g[off:off + rows * ld].view(rows, ld)[untouched].abs().sum()) == 0.0
        b1, b2, eps = 0.9, 0.999, 1e-8
        grad = g if decoupled else g + wd * arena
        if decoupled:
            arena.mul_(1 - lr * wd)
        m.lerp_(grad, 1 - b1)
        v.mul_(b2).addcmul_(grad, grad, value=1 - b2)
        denom = (v.sqrt() / math.sqrt(1 - b2 ** t)).add_(eps)
        arena.addcdiv_(m, denom, value=-lr / (1 - b1 ** t))
        g.zero_()


def _free_port():
    s = socket.socket(); s.bind(('127.0.0.1', 0)); p = s.getsockname()[1]; s.close()
    return p


def _worker(rank, world, port, kind, flags, out_dir, exchange='sparse', capq=None):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        from oracle import mf_oracle as O
        from hassaku_b200.sharded import ShardedMF, partition_batch_by_user_owner
        torch.set_num_threads(1)
        U, I, d, B, N = 53, 41, 6, 24, 5
        torch.manual_seed(7)
        ref = O.OracleMF(U, I, d, *flags)
        with torch.no_grad():
            for p in ref.parameters():
                p.copy_(torch.randn_like(p) * 0.3)
        lr, wd = 1e-2, 1e-3
        smf = ShardedMF(U, I, d, *flags, world=world, rank=rank, device='cpu', ops=TorchRefOps())
        smf.load_full_state_dict(ref.state_dict())
        tr = O.OracleTrainer(ref, kind, lr, wd, 'adamw', neg_train=N)
        rng = np.random.RandomState(3)
        shift = math.log(I / N) if kind == 'sampled_softmax' else 0.0
        losses_ref, losses = [], []
        for step in range(3):
            u = torch.from_numpy(rng.randint(0, U, B).astype(np.int64))
            i = torch.from_numpy(rng.randint(0, I, (B, N + 1)).astype(np.int64))
            if step == 1:
                i[:, 2] = i[:, 1]       # duplicate items inside rows, and across ranks
                u[:6] = u[0]
            losses_ref.append(float(tr.step(u, i)['loss']))
            ul, il = partition_batch_by_user_owner(u, i, world, rank)
            smf.step(ul, il, B, kind, shift, lr, wd, exchange=exchange, capq=capq)
            losses.append(smf.pop_loss())
        sd = smf.full_state_dict()
        if rank == 0:
            for n, p in ref.state_dict().items():
                a, b = sd[n].double(), p.double()
                assert a.shape == b.shape, n
                err = float((a - b).abs().max() / b.abs().max())
                assert err < 1e-5, (n, err)
            for a, b in zip(losses, losses_ref):
                assert abs(a - b) <= 1e-5 * abs(b), (losses, losses_ref)
            open(os.path.join(out_dir, 'ok'), 'w').write('ok')
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize('kind,flags', [('bpr', (False, True, False)), ('sampled_softmax', (False, True, False)),
                                        ('bce', (True, True, True))])
def test_sharded_train_step_world2_matches_oracle(kind, flags, tmp_path):
    port = _free_port()
    mp.spawn(_worker, args=(2, port, kind, flags, str(tmp_path)), nprocs=2, join=True)
    assert (tmp_path / 'ok').exists()


def test_sharded_dense_exchange_world2_matches_oracle(tmp_path):
    mp.spawn(_worker, args=(2, _free_port(), 'bce', (True, True, True), str(tmp_path), 'dense'), nprocs=2, join=True)
    assert (tmp_path / 'ok').exists()


@pytest.mark.parametrize('exchange,flags', [('dense', (True, True, True)), ('dense', (False, False, False)),
                                            ('sparse', (False, True, False))])
def test_sharded_step_world3_ragged_shards_matches_oracle(exchange, flags, tmp_path):
    """41 items / 53 users over 3 ranks: unequal shard sizes (14 / 14 / 13 items), bias block of the dense exchange
    rounded up to whole rows (cap 14, ld 8 -> 2 extra rows), and the dense exchange without any bias table."""
    mp.spawn(_worker, args=(3, _free_port(), 'bce', flags, str(tmp_path), exchange), nprocs=3, join=True)
    assert (tmp_path / 'ok').exists()


def test_exchange_capacity_bounds():
    from hassaku_b200.sharded import exchange_capacity
    assert exchange_capacity(8192 * 51, 1_000_000, 8) <= 125_000                  # cfg4: ~43 k expected per owner
    assert 40_000 < exchange_capacity(8192 * 51, 1_000_000, 8) < 50_000
    assert exchange_capacity(8192 * 51, 3706, 2) == 1853                          # cfg2: the whole shard
    assert exchange_capacity(10, 41, 3) == 14


def test_partition_and_shard_spec():
    from hassaku_b200.sharded import ShardSpec, partition_batch_by_user_owner
    s = ShardSpec(4, 1, 10, 7)
    assert s.n_local_users == len([u for u in range(10) if u % 4 == 1]) == 3
    assert s.n_local_items == len([i for i in range(7) if i % 4 == 1]) == 2
    u = torch.arange(10)
    i = torch.arange(20).view(10, 2)
    parts = [partition_batch_by_user_owner(u, i, 4, r) for r in range(4)]
    assert sum(len(p[0]) for p in parts) == 10
    assert all(((p[0] % 4) == r).all() for r, p in enumerate(parts))
