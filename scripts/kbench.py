"""Kernel micro-benchmarks (CUDA events, L2 flushed between launches) for A/B-ing kernel variants on the GPU box.
    python scripts/kbench.py train [--workload cfg2] [--variants regs,ring,q]
    python scripts/kbench.py trainraw --users 2000000 --items 1000000 --dim 128 --batch 8192 --variants auto
    python scripts/kbench.py eval --users 18944 --items 1000000 --dim 256 --batch 18944 --variants bf16,tf32
"""
import argparse
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def time_kernel(fn, iters=30, flush=None):
    ts = []
    for _ in range(iters):
        if flush is not None:
            flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ts = sorted(ts)
    return ts[len(ts) // 2], ts[0]


def train(args):
    from hassaku_b200 import _C
    from hassaku_b200.algorithms.sgd_alg import SGDMatrixFactorization
    from hassaku_b200.data.synthetic import make_named
    from hassaku_b200.train.optim import DenseAdam
    name, d, B, N, loss, lr, wd = bench.WORKLOADS[args.workload]
    data = make_named(name)
    U, I = data.n_users, data.n_items
    torch.manual_seed(64)
    model = SGDMatrixFactorization(U, I, d, use_item_bias=True).to('cuda')
    opt = DenseAdam(model, lr=lr, weight_decay=wd)
    us, its = bench.make_batches(data, B, N, 4)
    u = [torch.from_numpy(x).cuda() for x in us]
    i = [torch.from_numpy(x).cuda() for x in its]
    flush = torch.empty(256 << 20, dtype=torch.uint8, device='cuda')
    acc = torch.zeros(1, dtype=torch.float64, device='cuda')
    kind = _C.LOSS_KINDS[loss]
    shift = float(np.log(I / N)) if loss == 'sampled_softmax' else 0.0
    ab = bench.algorithmic_bytes(U, I, d, B, N)
    for v in args.variants.split(','):
        _C.TRAIN_VARIANT = v
        k = [0]

        def fn():
            _C.mf_train_fused(model._tables(), opt.grad_tables, u[k[0] % 4], i[k[0] % 4], kind, shift, acc)
            k[0] += 1
        for _ in range(3):
            fn()
        med, best = time_kernel(fn, 30, flush)
        medw, bestw = time_kernel(fn, 30, None)
        print(f'fused[{v:5s}] flushed med {med*1e3:8.1f} us best {best*1e3:8.1f} us | warm med {medw*1e3:8.1f} us  '
              f'-> {ab["gather_scatter"]/med/1e6:8.0f} GB/s algorithmic')
        opt.g.zero_()
    med, best = time_kernel(lambda: opt.step_fused(), 30, flush)
    print(f'adamw        flushed med {med*1e3:8.1f} us best {best*1e3:8.1f} us -> {ab["adamw"]/med/1e6:8.0f} GB/s (28 B/param)')


def trainraw(args):
    """fused step kernel alone on random tables / uniform indices of any shape (cfg4: --users 2000000 --items 1000000 --dim 128)"""
    from hassaku_b200 import _C
    from hassaku_b200.algorithms.sgd_alg import ArenaLayout
    U, I, d, B, N = args.users, args.items, args.dim, args.batch, args.neg
    lay = ArenaLayout(U, I, d, False, True, False)
    arena = torch.randn(lay.n_total, device='cuda') * 0.05
    g = torch.zeros_like(arena)
    gen = torch.Generator(device='cuda'); gen.manual_seed(0)
    u = [torch.randint(0, U, (B,), device='cuda', generator=gen) for _ in range(4)]
    i = [torch.randint(0, I, (B, N + 1), device='cuda', generator=gen) for _ in range(4)]
    flush = torch.empty(256 << 20, dtype=torch.uint8, device='cuda')
    acc = torch.zeros(1, dtype=torch.float64, device='cuda')
    kind = _C.LOSS_KINDS[args.loss]
    shift = float(np.log(I / N)) if args.loss == 'sampled_softmax' else 0.0
    ab = 4 * d * B * (N + 2)
    for v in args.variants.split(','):
        _C.TRAIN_VARIANT = v
        k = [0]

        def fn():
            _C.mf_train_fused(lay.tables(arena), lay.tables(g), u[k[0] % 4], i[k[0] % 4], kind, shift, acc)
            k[0] += 1
        for _ in range(3):
            fn()
        med, best = time_kernel(fn, 30, flush)
        print(f'fused[{v:5s}] {args.loss} U={U} I={I} d={d} B={B} N={N}: flushed med {med*1e3:8.1f} us best {best*1e3:8.1f} us'
              f' -> {ab/med/1e6:8.0f} GB/s algorithmic (rows read once)')
        g.zero_()


def evalk(args):
    """Evaluator scoring kernel alone: one user batch against the whole item table."""
    import math
    from hassaku_b200.algorithms.sgd_alg import SGDMatrixFactorization
    from hassaku_b200.eval.eval import DeviceCSR, TopKScorer
    from scipy import sparse as sp
    U, I, d, B = args.users, args.items, args.dim, args.batch
    torch.manual_seed(0)
    model = SGDMatrixFactorization(U, I, d, use_item_bias=True)
    with torch.no_grad():
        for p in model.parameters():
            p.copy_(torch.randn_like(p) * (1.0 / math.sqrt(d) if p.shape[-1] == d else 0.05))
    model.to('cuda')
    rng = np.random.RandomState(1)
    rows = np.repeat(np.arange(U), args.excl)
    ex = sp.csr_matrix((np.ones(len(rows), dtype=bool), (rows, rng.randint(0, I, len(rows)))), shape=(U, I))
    ex.sum_duplicates(); ex.sort_indices()
    ex = DeviceCSR(ex, 'cuda')
    users = torch.arange(B, device='cuda') % U
    flush = torch.empty(256 << 20, dtype=torch.uint8, device='cuda')
    for prec in args.variants.split(','):
        sc = TopKScorer(model, B, 100, prec, rescore=not args.no_rescore)
        fn = lambda: sc(users, ex)
        for _ in range(2):
            fn()
        med, best = time_kernel(fn, args.iters, flush)
        fl = 2.0 * B * I * d
        print(f'eval[{prec:4s}] U_b={B} I={I} d={d}: med {med:9.3f} ms best {best:9.3f} ms -> {B/med*1e3:12.0f} users/s '
              f'{fl/med/1e9:8.1f} TFLOP/s')


if __name__ == '__main__':
    ap = argparse.ArgumentParser()
    ap.add_argument('what', choices=['train', 'trainraw', 'eval'])
    ap.add_argument('--neg', type=int, default=50)
    ap.add_argument('--loss', default='bpr')
    ap.add_argument('--workload', default='cfg2')
    ap.add_argument('--variants', default='regs,ring,q')
    ap.add_argument('--users', type=int, default=6040)
    ap.add_argument('--items', type=int, default=3706)
    ap.add_argument('--dim', type=int, default=402)
    ap.add_argument('--batch', type=int, default=6040)
    ap.add_argument('--excl', type=int, default=80)
    ap.add_argument('--iters', type=int, default=10)
    ap.add_argument('--no-rescore', action='store_true', help='tensor-core modes: raw low-precision ranking (kernel alone)')
    a = ap.parse_args()
    {'train': train, 'trainraw': trainraw, 'eval': evalk}[a.what](a)
