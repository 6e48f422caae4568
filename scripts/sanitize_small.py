"""Small invocations of every kernel family for compute-sanitizer (memcheck / racecheck / synccheck) — shapes sized so that
the instrumented run finishes in a minute or two.
    compute-sanitizer --tool memcheck  python scripts/sanitize_small.py
    compute-sanitizer --tool racecheck python scripts/sanitize_small.py
"""
import math
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from hassaku_b200 import _C  # noqa: E402
from hassaku_b200.algorithms.sgd_alg import SGDMatrixFactorization  # noqa: E402
from hassaku_b200.eval.eval import DeviceCSR, TopKScorer  # noqa: E402
from hassaku_b200.sharded import ShardedMF  # noqa: E402
from hassaku_b200.train.optim import DenseAdam  # noqa: E402


def main():
    which = sys.argv[1:] or ['train', 'shard', 'eval']
    dev = 'cuda'
    torch.manual_seed(0)
    if 'train' in which:
        for d, variant, kind in ((402, 'ring', 0), (128, 'q', 1), (128, 'regs', 2), (64, 'q', 0)):
            U, I, B, N = 700, 500, 96, 12
            m = SGDMatrixFactorization(U, I, d, True, True, True).to(dev)
            opt = DenseAdam(m, lr=1e-3, weight_decay=1e-4)
            acc = torch.zeros(1, dtype=torch.float64, device=dev)
            u = torch.randint(0, U, (B,), device=dev); i = torch.randint(0, I, (B, N + 1), device=dev)
            i[:, 2] = i[:, 1]; u[:5] = u[0]
            for _ in range(2):
                _C.mf_train_fused(m._tables(), opt.grad_tables, u, i, kind, 0.5, acc, variant=variant)
                opt.mark(u, i)
                opt.step_fused()
            sc = torch.empty((B, N + 1), device=dev)
            _C.mf_scores(m._tables(), u, i, sc)
            m.check_status()
        # negative sampler
        indptr = torch.arange(0, 701 * 5, 5, dtype=torch.int64, device=dev)
        indices = torch.sort(torch.randint(0, 500, (700, 5), device=dev), dim=1).values.to(torch.int32).reshape(-1).contiguous()
        out = torch.empty((96, 13), dtype=torch.int64, device=dev)
        _C.sample_negatives(torch.randint(0, 700, (96,), device=dev), torch.randint(0, 500, (96,), device=dev), 12, 500, 700, indptr,
                            indices, 7, 3, out)
        print('train ok', flush=True)
    if 'shard' in which:
        U, I, d, B, N = 301, 2003, 128, 64, 10
        full = SGDMatrixFactorization(U, I, d, False, True, False)
        smf = ShardedMF(U, I, d, use_item_bias=True, world=1, rank=0, device=dev)
        smf.load_full_state_dict(full.state_dict())
        for ex in ('sparse', 'dense'):
            u = torch.randint(0, U, (B,), device=dev); i = torch.randint(0, I, (B, N + 1), device=dev)
            smf.step(u, i, B, 'bpr', 0.0, 1e-3, 1e-4, exchange=ex)
        smf.check_status()
        print('shard ok', flush=True)
    if 'eval' in which:
        from scipy import sparse as sp
        U, I, d, k = 300, 3001, 96, 100
        m = SGDMatrixFactorization(U, I, d, True, True, True)
        with torch.no_grad():
            for p in m.parameters():
                p.copy_(torch.randn_like(p) * (1 / math.sqrt(d) if p.shape[-1] == d else 0.1))
        m.to(dev)
        rng = np.random.RandomState(0)
        rows = np.repeat(np.arange(U), 20)
        ex = sp.csr_matrix((np.ones(len(rows), dtype=bool), (rows, rng.randint(0, I, len(rows)))), shape=(U, I))
        ex.sum_duplicates(); ex.sort_indices()
        exd = DeviceCSR(ex, dev)
        users = torch.arange(U, device=dev)
        for prec in ('fp32', 'bf16', 'tf32'):
            for variant in (('pair', 'single') if prec != 'fp32' else ('pair',)):
                _C.EVAL_TC_VARIANT = variant
                s, ids = TopKScorer(m, U, k, prec)(users, exd)
                torch.cuda.synchronize()
        m.check_status()
        print('eval ok', flush=True)


if __name__ == '__main__':
    main()
