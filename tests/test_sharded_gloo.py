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
    """Test-only stand-in for hassaku_b200.sharded.CudaOps."""

    def gather_rows(self, table2d, idx):
        return table2d[idx].contiguous()

    def local_index(self, idx, world, rank_stride):
        return (idx % world) * rank_stride + torch.div(idx, world, rounding_mode='floor')

    def scatter_add_rows(self, table2d, idx, rows):
        table2d.index_add_(0, idx, rows)

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

    def adamw(self, arena, m, v, g, lr, wd, t, decoupled=True):
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


def _worker(rank, world, port, kind, flags, out_dir, exchange='sparse', inplace=False):
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
        smf = ShardedMF(U, I, d, *flags, world=world, rank=rank, device='cpu', ops=TorchRefOps(), inplace_exchange=inplace)
        assert smf.inplace_exchange == inplace
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
            smf.step(ul, il, B, kind, shift, lr, wd, exchange=exchange)
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


@pytest.mark.parametrize('world,exchange,flags', [(2, 'dense', (True, True, True)), (3, 'dense', (False, False, False)),
                                                  (3, 'dense', (False, True, False)), (2, 'sparse', (False, True, False))])
def test_sharded_step_inplace_exchange_layout_matches_oracle(world, exchange, flags, tmp_path):
    """inplace_exchange=True: the item rows + biases live in the arena as the [capP, ld] block the collectives move
    (no staging copies); dense and sparse steps, state_dict round trip and AdamW over the padded block all still match
    the single-process oracle."""
    mp.spawn(_worker, args=(world, _free_port(), 'bce', flags, str(tmp_path), exchange, True), nprocs=world, join=True)
    assert (tmp_path / 'ok').exists()


def test_inplace_layout_offsets():
    from hassaku_b200.algorithms.sgd_alg import ArenaLayout
    lay = ArenaLayout(7, 5, 6, True, True, True, item_block_rows=8, item_bias_row=7)    # ld 8: 7 bias floats fit row 7
    a = torch.arange(lay.n_total, dtype=torch.float32)
    Uw, Vw, Ub, Ib, Gb = lay.views(a)
    assert Vw.shape == (5, 6) and Ib.shape == (5, 1)
    assert int(Ib[0, 0]) == lay.off_V + 7 * 8 and int(Vw[0, 0]) == lay.off_V
    assert lay.off_Ub >= lay.off_V + 8 * 8 and lay.off_Gb > lay.off_Ub          # nothing overlaps the item block
    plain = ArenaLayout(7, 5, 6, True, True, True)
    assert plain.off_Ib > plain.off_Ub and plain.n_total != lay.n_total


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
