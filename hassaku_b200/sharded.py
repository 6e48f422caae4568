"""Item-/user-sharded SGD-MF over the GPUs of one box: one process per GPU, torch.distributed (NCCL over
NVLink / NVSwitch) for the exchanges, the same sm_100a kernels for the arithmetic.  The reference's only multi-GPU
mechanism is nn.DataParallel (train/trainer.py:38-40), which re-broadcasts all parameters every forward and reduces
dense gradients to GPU 0; here nothing is replicated.

Partition (G = world size): item i lives on rank i % G as local row i // G (spreads the Zipf-popular low ids), with its
bias and AdamW state; user u likewise on rank u % G.  A training sample is processed by the owner of its USER, so user
rows, user AdamW state and the negative sampler are always local.

One training step (identical arithmetic to the single-GPU step on the union of the ranks' batches):
  1. dedupe the item ids of the local batch, group them by owner                       (index bookkeeping)
  2. all-to-all: ids to their owners                                                   (int64, ~unique ids)
  3. owners pack the requested rows (hsk_gather_rows) + biases, all-to-all back        (fp32 rows [cnt, ld])
  4. hsk_mf_train_fused_n on (local user shard, compact table of fetched rows) with GLOBAL normalisers
  5. all-to-all: compact row gradients to the owners, hsk_scatter_add_rows into the local dense gradient
  6. hsk_adamw_dense over the local arena (28 B / local parameter: the dominant term shards perfectly)
Evaluation: user rows of the batch are all-gathered, every rank scores them against its item shard (hsk_eval_topk with
id_offset = rank, id_stride = G), the per-shard top-k lists are exchanged (all-to-all) so that each rank merges
(hsk_topk_merge) and scores the metrics of ITS users; per-group sums are all-reduced once per sweep.

The index bookkeeping is device-agnostic torch code, the arithmetic goes through an `ops` object: `CudaOps` (the
kernels) in production; the world-size-2 gloo tests on CPU plug in a torch reference to check the routing.
"""
from __future__ import annotations

import math
import os
from typing import Dict, Optional

import numpy as np
import torch
import torch.distributed as dist

from hassaku_b200 import _C
from hassaku_b200.algorithms.sgd_alg import ArenaLayout, SGDMatrixFactorization


class ShardSpec:
    def __init__(self, world: int, rank: int, n_users: int, n_items: int):
        self.world, self.rank, self.n_users, self.n_items = world, rank, n_users, n_items
        self.n_local_users = len(range(rank, n_users, world))
        self.n_local_items = len(range(rank, n_items, world))

    def local_count(self, n: int, rank: int) -> int:
        return len(range(rank, n, self.world))


class CudaOps:
    """The arithmetic of the sharded step on the sm_100a kernels."""

    def __init__(self, status: Optional[torch.Tensor] = None):
        self.status = status

    def gather_rows(self, table2d: torch.Tensor, idx: torch.Tensor) -> torch.Tensor:
        out = torch.empty((idx.numel(), table2d.stride(0)), dtype=torch.float32, device=table2d.device)
        _C.gather_rows(table2d, idx, out, self.status)
        return out

    def scatter_add_rows(self, table2d: torch.Tensor, idx: torch.Tensor, rows: torch.Tensor):
        _C.scatter_add_rows(table2d, idx, rows.contiguous(), self.status)

    def local_index(self, idx: torch.Tensor, world: int, rank_stride: int) -> torch.Tensor:
        return _C.shard_local_index(idx.contiguous(), world, rank_stride)

    def train_fused(self, lay: ArenaLayout, arena, g_arena, Vc, Ibc, gVc, gIbc, u_local, compact_idx, B_global, kind, shift,
                    loss_accum):
        Uw, _, Ub, _, Gb = lay.views(arena)
        gU, _, gUb, _, gGb = lay.views(g_arena)
        t = _C.make_tables(Uw, Vc[:, :lay.d], Ub, Ibc, Gb, lay.d)
        g = _C.make_tables(gU, gVc[:, :lay.d], gUb, gIbc, gGb, lay.d)
        _C.mf_train_fused_n(t, g, u_local, compact_idx, B_global, kind, shift, loss_accum, self.status)

    def adamw(self, arena, m, v, g, lr, wd, t, decoupled=True):
        _C.adamw_dense(arena, m, v, g, lr, 0.9, 0.999, 1e-8, wd, t, arith=0, adam_l2=not decoupled, zero_grad=True)


def _a2a(inp: torch.Tensor, out_rows: int, in_splits, out_splits, group) -> torch.Tensor:
    out = torch.empty((out_rows,) + tuple(inp.shape[1:]), dtype=inp.dtype, device=inp.device)
    dist.all_to_all_single(out, inp.contiguous(), out_splits, in_splits, group=group)
    return out


class Exchange:
    """Steps 1-2 of the docstring for one batch: who needs which item rows."""

    def __init__(self, spec: ShardSpec, i_global: torch.Tensor, group=None):
        G = spec.world
        self.spec, self.group = spec, group
        flat = i_global.reshape(-1)
        uniq, inv = torch.unique(flat, sorted=True, return_inverse=True)
        owner = uniq % G
        order = torch.argsort(owner, stable=True)
        pos = torch.empty_like(order)
        pos[order] = torch.arange(order.numel(), device=order.device)
        self.n_uniq = int(uniq.numel())
        self.compact_idx = pos[inv].view(i_global.shape).contiguous()      # slot -> row of the compact table
        send_counts = torch.bincount(owner, minlength=G)
        recv_counts = torch.empty_like(send_counts)
        dist.all_to_all_single(recv_counts, send_counts, group=group)
        self.send_splits = [int(x) for x in send_counts.tolist()]
        self.recv_splits = [int(x) for x in recv_counts.tolist()]
        self.n_recv = sum(self.recv_splits)
        recv_ids = _a2a(uniq[order], self.n_recv, self.send_splits, self.recv_splits, group)
        self.recv_local_rows = torch.div(recv_ids, G, rounding_mode='floor')   # rows of MY shard peers asked for

    def fetch(self, rows_for_peers: torch.Tensor) -> torch.Tensor:
        """owner -> requester: rows [n_recv, w] in, compact table [n_uniq, w] out."""
        return _a2a(rows_for_peers, self.n_uniq, self.recv_splits, self.send_splits, self.group)

    def push(self, compact_rows: torch.Tensor) -> torch.Tensor:
        """requester -> owner: compact [n_uniq, w] in, [n_recv, w] out (aligned with recv_local_rows)."""
        return _a2a(compact_rows, self.n_recv, self.send_splits, self.recv_splits, self.group)


class ShardedMF:
    """The local shard of an SGDMatrixFactorization plus its optimizer state."""

    def __init__(self, n_users: int, n_items: int, d: int, use_user_bias=False, use_item_bias=False, use_global_bias=False,
                 world: Optional[int] = None, rank: Optional[int] = None, device='cuda', ops=None, group=None,
                 inplace_exchange: Optional[bool] = None):
        """`inplace_exchange` (default: env HSK_SHARDED_INPLACE == '1', else off; opt-in until measured on GPUs): lay the
        item rows and item biases out in the arena as the [capP, ld] block the dense exchange sends, so that the
        all-gather reads the arena and the reduce-scatter writes the gradient arena directly (4 small copies fewer
        per step)."""
        world = dist.get_world_size(group) if world is None else world
        rank = dist.get_rank(group) if rank is None else rank
        self.spec = ShardSpec(world, rank, n_users, n_items)
        self.d, self.group, self.device = d, group, torch.device(device)
        self.flags = (use_user_bias, use_item_bias, use_global_bias)
        if inplace_exchange is None:
            inplace_exchange = os.environ.get('HSK_SHARDED_INPLACE') == '1'
        self.inplace_exchange = bool(inplace_exchange)
        if self.inplace_exchange:
            ld = (d + 3) // 4 * 4
            cap = math.ceil(n_items / world)
            self.layout = ArenaLayout(self.spec.n_local_users, self.spec.n_local_items, d, *self.flags,
                                      item_block_rows=cap + math.ceil(cap / ld), item_bias_row=cap)
        else:
            self.layout = ArenaLayout(self.spec.n_local_users, self.spec.n_local_items, d, *self.flags)
        self.arena = torch.zeros(self.layout.n_total, dtype=torch.float32, device=self.device)
        self.m = torch.zeros_like(self.arena)
        self.v = torch.zeros_like(self.arena)
        self.g = torch.zeros_like(self.arena)
        self.t = 0
        self.status = torch.zeros(1, dtype=torch.int32, device=self.device) if self.device.type == 'cuda' else None
        self.ops = ops if ops is not None else CudaOps(self.status)
        self.loss_accum = torch.zeros(1, dtype=torch.float64, device=self.device)

    # ---- state in / out ----
    def load_full_state_dict(self, sd: Dict[str, torch.Tensor]):
        """Take this rank's rows of a full (reference-format) state_dict."""
        G, r = self.spec.world, self.spec.rank
        Uw, Vw, Ub, Ib, Gb = self.layout.views(self.arena)
        with torch.no_grad():
            Uw.copy_(sd['user_embeddings.weight'][r::G].to(self.device))
            Vw.copy_(sd['item_embeddings.weight'][r::G].to(self.device))
            if Ub is not None:
                Ub.copy_(sd['user_bias.weight'][r::G].to(self.device))
            if Ib is not None:
                Ib.copy_(sd['item_bias.weight'][r::G].to(self.device))
            if Gb is not None:
                Gb.copy_(sd['global_bias'].to(self.device))

    def full_state_dict(self) -> Dict[str, torch.Tensor]:
        """All-gather the shards into a full state_dict (checkpoint in the reference's format; parity tests)."""
        G = self.spec.world
        Uw, Vw, Ub, Ib, Gb = self.layout.views(self.arena)

        def gather(local, n_total):
            width = local.shape[1]
            cap = math.ceil(n_total / G)
            buf = torch.zeros((cap, width), dtype=torch.float32, device=self.device)
            buf[:local.shape[0]] = local
            parts = [torch.empty_like(buf) for _ in range(G)]
            dist.all_gather(parts, buf, group=self.group)
            full = torch.empty((n_total, width), dtype=torch.float32, device=self.device)
            for q in range(G):
                full[q::G] = parts[q][:self.spec.local_count(n_total, q)]
            return full.cpu()

        sd = {'user_embeddings.weight': gather(Uw, self.spec.n_users), 'item_embeddings.weight': gather(Vw, self.spec.n_items)}
        if Ub is not None:
            sd['user_bias.weight'] = gather(Ub, self.spec.n_users)
        if Ib is not None:
            sd['item_bias.weight'] = gather(Ib, self.spec.n_items)
        if Gb is not None:
            sd['global_bias'] = Gb.detach().cpu().clone()
        return sd

    # ---- one training step ----
    def train_step(self, u_global: torch.Tensor, i_global: torch.Tensor, B_global: int, loss_kind: str, neg_shift: float,
                   lr: float, wd: float, decoupled: bool = True):
        """u_global int64 [B_r] (all owned by this rank: u % G == rank), i_global int64 [B_r, 1+N] global item ids."""
        G, lay = self.spec.world, self.layout
        ld = lay.ld
        ex = Exchange(self.spec, i_global, self.group)
        _, Vw, _, Ib, _ = lay.views(self.arena)
        V2d = self.arena[lay.off_V:lay.off_V + lay.n_items * ld].view(lay.n_items, ld)
        gV2d = self.g[lay.off_V:lay.off_V + lay.n_items * ld].view(lay.n_items, ld)
        # 3. owners pack rows (+ bias) and send them back
        Vc = ex.fetch(self.ops.gather_rows(V2d, ex.recv_local_rows))
        Ibc = None
        if Ib is not None:
            Ibc = ex.fetch(Ib.view(-1)[ex.recv_local_rows].contiguous().view(-1, 1)).view(-1)
        # 4. local compute on (user shard, compact item table)
        gVc = torch.zeros_like(Vc)
        gIbc = torch.zeros_like(Ibc) if Ibc is not None else None
        u_local = torch.div(u_global, G, rounding_mode='floor')
        self.ops.train_fused(lay, self.arena, self.g, Vc, Ibc, gVc, gIbc, u_local, ex.compact_idx, B_global,
                             _C.LOSS_KINDS[loss_kind], neg_shift, self.loss_accum)
        # 5. gradients home
        self.ops.scatter_add_rows(gV2d, ex.recv_local_rows, ex.push(gVc))
        if gIbc is not None:
            gIb = lay.views(self.g)[3].view(-1)
            gIb.index_add_(0, ex.recv_local_rows, ex.push(gIbc.view(-1, 1)).view(-1))
        gGb = lay.views(self.g)[4]
        if gGb is not None:
            dist.all_reduce(gGb, group=self.group)   # the global bias is replicated: every rank applies the summed gradient
        # 6. optimizer on the local shard
        self.t += 1
        self.ops.adamw(self.arena, self.m, self.v, self.g, lr, wd, self.t, decoupled)

    # ---- dense exchange: when the batch touches (nearly) every item anyway ----
    def _dense_buffers(self):
        """Send / replica buffers of the dense exchange.  One rank's block is [capP, ld] floats: rows [0, cap) are its item
        rows (cap = ceil(n_items / G), unused rows zero), rows [cap, capP) hold its cap item biases flat - so ONE
        all-gather moves rows and biases, and ONE reduce-scatter brings back both gradients.  The fused kernel indexes
        the replica as a [G * capP, ld] table (row of item i = (i % G) * capP + i // G) and needs the biases under the
        same index, hence the compact `Ib` / `gIb` vectors of G * capP floats filled / drained by one strided copy."""
        if getattr(self, '_dense', None) is None:
            G, lay = self.spec.world, self.layout
            cap = math.ceil(self.spec.n_items / G)
            capP = cap + math.ceil(cap / lay.ld)
            z = lambda *shape: torch.zeros(shape, dtype=torch.float32, device=self.device)
            grads = z(G * capP * lay.ld + G * capP)          # [gV replica | gIb compact]: one memset per step
            self._dense = {'cap': cap, 'capP': capP, 'V': z(G * capP, lay.ld), 'Ib': z(G * capP),
                           'grads': grads, 'gV': grads[:G * capP * lay.ld].view(G * capP, lay.ld),
                           'gIb': grads[G * capP * lay.ld:]}
            if self.inplace_exchange:      # the arena / gradient arena hold the block themselves
                blk = slice(lay.off_V, lay.off_V + capP * lay.ld)
                self._dense['send'] = self.arena[blk].view(capP, lay.ld)
                self._dense['recv'] = self.g[blk].view(capP, lay.ld)
            else:
                self._dense['send'], self._dense['recv'] = z(capP, lay.ld), z(capP, lay.ld)
        return self._dense

    def _reduce_scatter(self, out: torch.Tensor, inp: torch.Tensor):
        if dist.get_backend(self.group) == 'nccl':
            dist.reduce_scatter_tensor(out, inp, group=self.group)
        else:  # gloo (CPU tests) has no reduce-scatter
            dist.all_reduce(inp, group=self.group)
            n = out.numel()
            out.copy_(inp.view(-1)[self.spec.rank * n:(self.spec.rank + 1) * n].view_as(out))

    def _dense_exchange_and_fused(self, u_global, i_global, B_global, loss_kind, neg_shift):
        """All-gather of the item shards into the rank-major replica, the ordinary fused kernel on it, reduce-scatter of
        the dense item gradient to the owners.  No dedupe, no host sync, fixed shapes (CUDA-graph friendly): 2 collectives,
        1 memset, 2 index kernels, 4 small copies around the fused kernel."""
        G, lay = self.spec.world, self.layout
        ld, nl = lay.ld, lay.n_items
        D = self._dense_buffers()
        cap, capP = D['cap'], D['capP']
        _, _, _, Ib, _ = lay.views(self.arena)
        has_ib = Ib is not None
        if not self.inplace_exchange:
            D['send'][:nl] = self.arena[lay.off_V:lay.off_V + nl * ld].view(nl, ld)
            if has_ib:
                D['send'].view(-1)[cap * ld:cap * ld + nl] = Ib.view(-1)
        dist.all_gather_into_tensor(D['V'], D['send'], group=self.group)
        if has_ib:
            D['Ib'].view(G, capP)[:, :cap] = D['V'].view(G, capP * ld)[:, cap * ld:cap * ld + cap]
        rows = self.ops.local_index(i_global, G, capP)
        u_local = self.ops.local_index(u_global, G, 0)
        D['grads'].zero_()
        self.ops.train_fused(lay, self.arena, self.g, D['V'], D['Ib'] if has_ib else None, D['gV'],
                             D['gIb'] if has_ib else None, u_local, rows, B_global,
                             _C.LOSS_KINDS[loss_kind], neg_shift, self.loss_accum)
        if has_ib:
            D['gV'].view(G, capP * ld)[:, cap * ld:cap * ld + cap] = D['gIb'].view(G, capP)[:, :cap]
        # in-place: the block of the gradient arena is otherwise untouched by a dense step (item gradients went to the
        # replica), so the reduce-scatter writes it directly
        self._reduce_scatter(D['recv'], D['gV'])
        if not self.inplace_exchange:
            self.g[lay.off_V:lay.off_V + nl * ld].view(nl, ld).add_(D['recv'][:nl])
            if has_ib:
                lay.views(self.g)[3].view(-1).add_(D['recv'].view(-1)[cap * ld:cap * ld + nl])
        gGb = lay.views(self.g)[4]
        if gGb is not None:
            dist.all_reduce(gGb, group=self.group)

    def train_step_dense(self, u_global: torch.Tensor, i_global: torch.Tensor, B_global: int, loss_kind: str,
                         neg_shift: float, lr: float, wd: float, decoupled: bool = True):
        """Same step with a DENSE exchange (see _dense_exchange_and_fused).  The right choice when B (N + 1) >> n_items
        (cfg2: 418 k slots on 3 706 items, every row is requested by every rank each step anyway)."""
        self._dense_exchange_and_fused(u_global, i_global, B_global, loss_kind, neg_shift)
        self.t += 1
        self.ops.adamw(self.arena, self.m, self.v, self.g, lr, wd, self.t, decoupled)

    # ---- the dense step as ONE CUDA graph (fixed shapes): removes the launches / collectives worth of host latency ----
    CONST_TABLE_STEPS = 32768   # 1 MB table, 0.1 s to fill: a refill (host sync) practically never lands inside a timed region

    def _dense_body(self, u_global, i_global, B_global, loss_kind, neg_shift, consts_dev, decoupled):
        """train_step_dense with the AdamW scalars read from device memory (capturable)."""
        self._dense_exchange_and_fused(u_global, i_global, B_global, loss_kind, neg_shift)
        _C.adamw_dense_graph(self.arena, self.m, self.v, self.g, consts_dev, decoupled=decoupled, adam_l2=not decoupled)

    def _fill_const_table(self, gs, lr, wd):
        host = torch.empty((self.CONST_TABLE_STEPS, 8), dtype=torch.float32).pin_memory()
        for j in range(self.CONST_TABLE_STEPS):
            _C.adamw_consts(lr, 0.9, 0.999, 1e-8, wd, self.t + 1 + j, host[j])
        gs['table'].copy_(host)
        gs['step_idx'].zero_()
        gs['table_base_t'] = self.t

    def train_step_dense_graphed(self, u_global, i_global, B_global, loss_kind, neg_shift, lr, wd, decoupled=True):
        """train_step_dense captured once per (batch shape, hyper-parameters) and replayed.  The per-step AdamW scalars
        come from a device table of the next CONST_TABLE_STEPS steps indexed by a device-side step counter, so a replay
        needs no host-computed kernel argument."""
        key = (tuple(i_global.shape), B_global, loss_kind, float(neg_shift), float(lr), float(wd), bool(decoupled))
        gs = getattr(self, '_graph_state', None)
        if gs is None or gs['key'] != key:
            dev = self.device
            gs = {'key': key, 'u': torch.empty_like(u_global, device=dev), 'i': torch.empty_like(i_global, device=dev),
                  'table': torch.empty((self.CONST_TABLE_STEPS, 8), dtype=torch.float32, device=dev),
                  'consts': torch.empty(8, dtype=torch.float32, device=dev),
                  'step_idx': torch.zeros(1, dtype=torch.int64, device=dev)}
            self._fill_const_table(gs, lr, wd)
            # NCCL channels for these collectives must exist before capture: warm them up on the gradient buffers
            D = self._dense_buffers()
            dist.all_gather_into_tensor(D['gV'], D['recv'], group=self.group)
            self._reduce_scatter(D['recv'], D['gV'])
            torch.cuda.synchronize()
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                torch.index_select(gs['table'], 0, gs['step_idx'], out=gs['consts'].view(1, 8))
                self._dense_body(gs['u'], gs['i'], B_global, loss_kind, neg_shift, gs['consts'], decoupled)
                gs['step_idx'].add_(1)
            gs['graph'] = graph
            self._graph_state = gs
        if self.t - gs['table_base_t'] >= self.CONST_TABLE_STEPS:
            torch.cuda.current_stream().synchronize()
            self._fill_const_table(gs, lr, wd)
        gs['u'].copy_(u_global, non_blocking=True)
        gs['i'].copy_(i_global, non_blocking=True)
        gs['graph'].replay()
        self.t += 1

    def close(self):
        """Drop the captured CUDA graphs (they hold NCCL work): call before dist.destroy_process_group(), which
        otherwise blocks."""
        torch.cuda.synchronize()
        self._graph_state = None
        import gc
        gc.collect()
        torch.cuda.synchronize()

    def step(self, u_global, i_global, B_global, loss_kind, neg_shift, lr, wd, decoupled=True, exchange='auto'):
        """Dispatch on the expected fraction of distinct items: dense exchange when the batch covers the item table."""
        if exchange == 'auto':
            exchange = 'dense' if i_global.numel() >= 2 * self.spec.n_items // self.spec.world else 'sparse'
        fn = {'dense': self.train_step_dense, 'dense_graph': self.train_step_dense_graphed, 'sparse': self.train_step}[exchange]
        return fn(u_global, i_global, B_global, loss_kind, neg_shift, lr, wd, decoupled)

    def pop_loss(self) -> float:
        """Sum over ranks of the batch-mean loss contributions since the last call (one all-reduce + host sync)."""
        t = self.loss_accum.clone()
        dist.all_reduce(t, group=self.group)
        self.loss_accum.zero_()
        return float(t.item())

    def _csr_cache(self, m):
        from hassaku_b200.eval.eval import DeviceCSR
        cache = self.__dict__.setdefault('_csr', {})
        if id(m) not in cache:
            cache[id(m)] = (m, DeviceCSR(m, self.device))   # keep `m` alive so the id stays unique
        return cache[id(m)][1]

    # ---- item-sharded full-rank evaluation ----
    def evaluate(self, labels_csr, exclude_csr, evaluator, batch_size: int = 8192, precision: str = 'fp32'):
        """Every rank evaluates ITS users (u % G == rank) against ALL items; returns the global metric dict on every
        rank.  `labels_csr` / `exclude_csr`: full scipy CSR matrices (global ids)."""
        from hassaku_b200.eval.eval import DeviceCSR
        if precision not in _C.PRECISIONS:
            raise ValueError(f'eval precision {precision!r} not in {sorted(_C.PRECISIONS)}')
        prec = _C.PRECISIONS[precision]          # 0: fp32-exact SIMT kernel; tf32 / bf16: tcgen05 kernel on the local shard
        G, r, lay, dev = self.spec.world, self.spec.rank, self.layout, self.device
        k = max(evaluator.K_VALUES)
        labels = labels_csr if isinstance(labels_csr, DeviceCSR) else self._csr_cache(labels_csr)
        exclude = exclude_csr if isinstance(exclude_csr, DeviceCSR) else self._csr_cache(exclude_csr)
        cap = math.ceil(self.spec.n_users / G)
        bs = min(batch_size, cap)
        ld = lay.ld
        U2d = self.arena[lay.off_U:lay.off_U + lay.n_users * ld].view(lay.n_users, ld)
        Uw, Vw, Ub, Ib, Gb = lay.views(self.arena)
        Be = G * bs
        if prec == 0:
            scratch = torch.empty(_C.eval_topk_scratch_bytes(Be, lay.n_items, k), dtype=torch.uint8, device=dev)
        else:
            scratch = torch.empty(_C.eval_topk_tc_scratch_bytes(Be, lay.n_items, k), dtype=torch.uint8, device=dev)
            Vq = _C.pack_rows(Vw.detach(), lay.d, prec)         # the local item shard, packed once per sweep
        top_s = torch.empty((Be, k), dtype=torch.float32, device=dev)
        top_i = torch.empty((Be, k), dtype=torch.int32, device=dev)
        u_rows = torch.arange(Be, dtype=torch.int64, device=dev)
        for s in range(0, cap, bs):
            n_mine = max(0, min(bs, lay.n_users - s))                 # my users of this round (may be fewer at the tail)
            rows = torch.zeros((bs, ld), dtype=torch.float32, device=dev)
            ubias = torch.zeros(bs, dtype=torch.float32, device=dev)
            gids = torch.full((bs,), -1, dtype=torch.int64, device=dev)
            if n_mine > 0:
                rows[:n_mine] = U2d[s:s + n_mine]
                gids[:n_mine] = (torch.arange(s, s + n_mine, device=dev) * G + r)
                if Ub is not None:
                    ubias[:n_mine] = Ub.view(-1)[s:s + n_mine]
            all_rows = torch.empty((Be, ld), dtype=torch.float32, device=dev)
            all_gids = torch.empty(Be, dtype=torch.int64, device=dev)
            all_ub = torch.empty(Be, dtype=torch.float32, device=dev)
            dist.all_gather_into_tensor(all_rows, rows, group=self.group)
            dist.all_gather_into_tensor(all_gids, gids, group=self.group)
            dist.all_gather_into_tensor(all_ub, ubias, group=self.group)
            valid = all_gids >= 0
            safe_gids = torch.where(valid, all_gids, torch.zeros_like(all_gids))
            if prec == 0:
                t = _C.make_tables(all_rows[:, :lay.d], Vw, all_ub if Ub is not None else None, Ib, Gb, lay.d)
                _C.eval_topk(t, safe_gids, k, top_s, top_i, scratch, exclude.indptr, exclude.indices, id_offset=r,
                             id_stride=G, status=self.status, u_rows=u_rows, n_users_global=self.spec.n_users)
            else:
                Uq = _C.pack_rows(all_rows[:, :lay.d], lay.d, prec)
                _C.eval_topk_tc(Uq, Vq, prec, safe_gids, self.spec.n_users, k, top_s, top_i, scratch,
                                Ub=all_ub if Ub is not None else None, Ib=Ib, Gb=Gb, excl_indptr=exclude.indptr,
                                excl_indices=exclude.indices, id_offset=r, id_stride=G, status=self.status, u_rows=u_rows)
            # exchange: rank q receives the G partial lists of its bs users
            recv_s = torch.empty((G, bs, k), dtype=torch.float32, device=dev)
            recv_i = torch.empty((G, bs, k), dtype=torch.int32, device=dev)
            dist.all_to_all_single(recv_s, top_s.view(G, bs, k), group=self.group)
            dist.all_to_all_single(recv_i, top_i.view(G, bs, k), group=self.group)
            if n_mine > 0:
                ms = torch.empty((bs, k), dtype=torch.float32, device=dev)
                mi = torch.empty((bs, k), dtype=torch.int32, device=dev)
                _C.topk_merge(recv_s, recv_i, ms, mi)
                evaluator.eval_batch_topk(gids[:n_mine].contiguous(), mi[:n_mine].contiguous(), labels)
        # one all-reduce of the accumulators per sweep
        evaluator._prepare(dev)
        dist.all_reduce(evaluator._sums, group=self.group)
        dist.all_reduce(evaluator._counts, group=self.group)
        if getattr(evaluator, '_hit_sums', None) is not None:
            dist.all_reduce(evaluator._hit_sums, group=self.group)
        return evaluator.get_results()


def partition_batch_by_user_owner(u_idxs: torch.Tensor, i_idxs: torch.Tensor, world: int, rank: int):
    """The rows of a GLOBAL batch this rank processes (sample -> owner of its user)."""
    sel = (u_idxs % world) == rank
    return u_idxs[sel].contiguous(), i_idxs[sel].contiguous()
