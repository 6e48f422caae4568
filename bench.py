"""bench.py — BPR-MF train triples/s (+ full-rank eval users/s, NDCG@10) on B200, the metric of BASELINE.json.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload cfg2|cfg1|cfg3]

Workload at N = 1: BASELINE.json configs[1] ("cfg2"): synthetic ML-1M shape (6 040 users x 3 706 items, ~0.93 M unique
interactions, 80/10/10 split), d = 402, train_batch_size 8192, neg_train 50, item bias, BPR, AdamW(lr 3e-4, wd 4e-5),
fp32 exact mode, followed by one full-rank evaluation sweep over all users.

A "step" = one pass of the hot path over one batch: hsk_mf_train_fused (gather + score + BPR loss + gradient scatter)
+ hsk_adamw_dense (dense AdamW over all parameters, gradient zeroing fused).  Prints ONE JSON line (rank 0).

  value      whole-job triples/s with the batches already resident in HBM, device-timed (CUDA events, max over ranks);
             L2 is flushed between timed steps (the 63 MB of tables + optimizer state fit the 126 MB L2, so without the
             flush the number is an L2-bandwidth number; it is reported as `value_l2_warm`)
  e2e        the same metric through the public API (FusedMFTrainStep called with HOST batches): per step the H2D copy
             of the int64 index batch from pinned memory and a D2H read of the loss are inside the timed region
  roofline   dominant kernel: algorithmic bytes per launch / mean launch duration (CUDA events on the launch stream)
             against the measured HBM copy bandwidth of MEASURED_PEAKS.json
  cpu_baseline / --impl reference   the oracle port of the reference's torch path (oracle/mf_oracle.py: the same ATen
             op sequence as train/trainer.py:133-148) timed on this box's host cores on a bounded sample
"""
import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

WORKLOADS = {
    #        data     d    B     N    loss               lr     wd
    'cfg1': ('ml1m', 402, 128, 50, 'bpr', 3e-4, 4e-5),
    'cfg2': ('ml1m', 402, 8192, 50, 'bpr', 3e-4, 4e-5),
    'cfg3': ('ml10m', 128, 8192, 100, 'sampled_softmax', 3e-4, 4e-5),
}


def algorithmic_bytes(U, I, d, B, N, item_bias=True):
    """A = 4 d B (N+2) (every gathered row once); small = index + bias traffic; 28 B per parameter."""
    A = 4 * d * B * (N + 2)
    small = 8 * B * (N + 1) + 8 * B * (N + 2)
    P = (U + I) * d + (I if item_bias else 0)
    return {'gather_scatter': 2 * A + small, 'adamw': 28 * P, 'total': 2 * A + small + 28 * P, 'P': P, 'A': A}


def measured_peaks():
    try:
        with open(os.path.join(ROOT, 'MEASURED_PEAKS.json')) as f:
            p = json.load(f)
        return float(p['hbm_gbs']), 'measured (MEASURED_PEAKS.json)'
    except Exception:
        return 6650.0, 'fallback (B200_PROFILING.md)'


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ('clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,'
         'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
         'clocks_event_reasons.sw_power_cap')

    def __init__(self, gpu_index=0, enabled=True):
        """gpu_index: one index or a comma-separated list (N > 1: rank 0 samples every GPU of the job; one nvidia-smi
        per rank would put N processes on the driver's global lock inside the timed region)."""
        self.rows, self.proc, self.gpu, self.enabled = [], None, gpu_index, enabled

    def wait_ready(self, timeout=8.0):
        """Block until the first sample arrived: nvidia-smi's start-up (NVML init over every GPU of the box, hundreds of
        ms, serialised on a driver lock) must be over BEFORE the timed region starts."""
        t0 = time.time()
        while self.proc is not None and not self.rows and self.proc.poll() is None and time.time() - t0 < timeout:
            time.sleep(0.01)

    def __enter__(self):
        if not self.enabled:
            return self
        try:
            self.proc = subprocess.Popen(['nvidia-smi', f'--query-gpu={self.Q}', '--format=csv,noheader,nounits',
                                          '-i', str(self.gpu), '-lms', '100'], stdout=subprocess.PIPE, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(',')])

    def __exit__(self, *a):
        if self.proc is not None:
            time.sleep(0.15)
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                pass

    def summary(self):
        sm, mx, reasons = [], [], set()
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
                for n, v in zip(names, r[3:7]):
                    if v.lower().startswith('active'):
                        reasons.add(n)
            except Exception:
                pass
        if not sm:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': [], 'samples': 0}
        top = sorted(sm)[len(sm) // 2:]  # samples under load = upper half
        return {'sm_mhz': float(np.median(top)), 'sm_max_mhz': max(mx), 'reasons': sorted(reasons), 'samples': len(sm)}


def make_batches(data, B, N, n_batches, seed=64):
    """Captured training batches with the reference loader's semantics (data/dataloader.py:92-129): positives drawn
    from the shuffled train interactions, N uniform negatives per row none of which is a train item of the user."""
    rng = np.random.default_rng(seed)
    coo = data.train.tocoo()
    I = data.n_items
    keys = np.sort(coo.row.astype(np.int64) * I + coo.col.astype(np.int64))
    us, its = [], []
    for _ in range(n_batches):
        sel = rng.integers(0, coo.nnz, B)
        u = coo.row[sel].astype(np.int64)
        neg = rng.integers(0, I, (B, N), dtype=np.int64)
        while True:
            k = (u[:, None] * I + neg).ravel()
            pos = np.searchsorted(keys, k)
            hit = (keys[np.minimum(pos, len(keys) - 1)] == k).reshape(B, N)
            if not hit.any():
                break
            neg[hit] = rng.integers(0, I, int(hit.sum()), dtype=np.int64)
        us.append(u)
        its.append(np.column_stack([coo.col[sel].astype(np.int64), neg]))
    return us, its


def init_like_reference(model, seed=64):
    import torch
    torch.manual_seed(seed)  # conf_parser.py:18 default seed; weights ~ N(0, (0.1/shape[-1])^2) (train/utils.py:13)
    return model


# ------------------------------------------------------------------------------------------------
def run_reference(args, wl):
    """--impl reference: the reference's CPU implementation of the step (oracle port) on all host threads."""
    import torch
    from oracle import mf_oracle as O
    from hassaku_b200.data.synthetic import make_named
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    name, d, B, N, loss, lr, wd = wl
    data = make_named(name)
    U, I = data.n_users, data.n_items
    torch.manual_seed(64)
    model = O.OracleMF(U, I, d, use_item_bias=True)
    tr = O.OracleTrainer(model, loss, lr, wd, 'adamw', neg_train=N)
    nb = min(args.steps + args.warmup, 8)
    us, its = make_batches(data, B, N, nb)
    us = [torch.from_numpy(x) for x in us]
    its = [torch.from_numpy(x) for x in its]
    labels = O.make_labels(B, N + 1)
    for s in range(args.warmup):
        tr.step(us[s % nb], its[s % nb], labels)
    t0 = time.perf_counter()
    for s in range(args.steps):
        tr.step(us[s % nb], its[s % nb], labels)
    dt = time.perf_counter() - t0
    v = args.steps * B * N / dt
    cores = torch.get_num_threads()
    line = {'impl': 'reference', 'metric': 'BPR-MF train triples/s', 'value': v, 'unit': 'triples/s', 'n_gpus': args.gpus,
            'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': 1e3 * dt / args.steps, 'higher_is_better': True,
            'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
            'config': workload_config(args.workload, wl, U, I),
            'cpu_baseline': {'value': v, 'unit': 'triples/s', 'cores': cores, 'kind': 'port',
                             'sample': f'{args.steps} steps of {args.workload} (B={B}, N={N}) on the oracle port of '
                                       f'train/trainer.py:133-148 (torch CPU, {cores} threads, {os.cpu_count()} cpus)'},
            'e2e': {'value': v, 'unit': 'triples/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}}
    print(json.dumps(line))


def workload_config(wname, wl, U, I):
    name, d, B, N, loss, lr, wd = wl
    return {'workload': f'{wname}: BPR-MF train step + full-rank eval, synthetic {name} shape', 'n_users': U, 'n_items': I,
            'embedding_dim': d, 'train_batch_size': B, 'neg_train': N, 'rec_loss': loss, 'optimizer': 'adamw', 'lr': lr,
            'wd': wd, 'item_bias': True, 'eval_batch_size': 8192, 'precision': 'fp32 exact',
            'l2': 'flushed between timed steps (tables + optimizer state fit L2); value_l2_warm = back to back'}


def cpu_eval_baseline(wl, data, n_users=512, eval_batch_size=256):
    """The reference's full-rank evaluation (eval/eval.py:237-253: [Be, I, d] broadcast product, dense mask, torch.topk,
    12 metrics per batch; eval batch 256 as in its README) on a bounded sample of users, host cores only."""
    import torch
    from oracle import mf_oracle as O
    name, d, B, N, loss, lr, wd = wl
    torch.manual_seed(64)
    model = O.OracleMF(data.n_users, data.n_items, d, use_item_bias=True)
    users = np.arange(min(n_users, data.n_users))
    t0 = time.perf_counter()
    with torch.no_grad():
        res = O.evaluate(model, data.val.tocsr(), data.train.tocsr(), eval_batch_size=eval_batch_size, users=users)
    dt = time.perf_counter() - t0
    cores = torch.get_num_threads()
    return {'value': len(users) / dt, 'unit': 'users/s', 'cores': cores, 'kind': 'port',
            'sample': f'{len(users)} of {data.n_users} users against all {data.n_items} items, eval batch {eval_batch_size}, '
                      f'oracle port of eval/eval.py:237-253 (torch CPU, {cores} threads), {dt:.1f} s',
            'ndcg@10_of_sample': float(res['ndcg@10'])}


def cpu_baseline(wl, data, us, its, budget_s=20.0):
    import torch
    from oracle import mf_oracle as O
    name, d, B, N, loss, lr, wd = wl
    torch.manual_seed(64)
    model = O.OracleMF(data.n_users, data.n_items, d, use_item_bias=True)
    tr = O.OracleTrainer(model, loss, lr, wd, 'adamw', neg_train=N)
    labels = O.make_labels(B, N + 1)
    tr.step(torch.from_numpy(us[0]), torch.from_numpy(its[0]), labels)  # warm-up
    n, t0 = 0, time.perf_counter()
    while True:
        tr.step(torch.from_numpy(us[n % len(us)]), torch.from_numpy(its[n % len(us)]), labels)
        n += 1
        dt = time.perf_counter() - t0
        if dt > budget_s or n >= 30:
            break
    cores = torch.get_num_threads()
    return {'value': n * B * N / dt, 'unit': 'triples/s', 'cores': cores, 'kind': 'port',
            'sample': f'{n} steps of the same workload (B={B}, N={N}) on the oracle port (torch CPU, {cores} threads of '
                      f'{os.cpu_count()} cpus), {dt:.1f} s'}


def run_extras(dev, flush):
    """Kernel-level measurements at the shapes of BASELINE configs 4 and 5 that do not fit the cfg2 step (reported next to
    the headline, never instead of it):
      adamw_cfg4      hsk_adamw_dense over P = 385 M parameters (2 M users x 1 M items, d 128): the genuinely HBM-bound
                      kernel of the path (28 B / parameter + 4 B gradient zeroing), GB/s vs the measured copy bandwidth
      eval_tc_cfg5    hsk_eval_topk_tc, one batch of 18 944 users (148 CTAs x 128) against 1 M items, d 256, BF16 and
                      TF32: users/s, TFLOP/s vs the measured sustained bf16 GEMM peak (tensor-pipe roofline)"""
    import torch
    from hassaku_b200 import _C
    out = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))
    except Exception:
        peaks = {'hbm_gbs': 6650.0, 'bf16_tflops_sustained': 1400.0, 'bf16_tflops': 1590.0}

    def timed(fn, iters):
        ts = []
        for _ in range(iters):
            flush.zero_()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); fn(); b.record()
            torch.cuda.synchronize()
            ts.append(a.elapsed_time(b))
        return float(np.median(ts))

    try:
        P = (2_000_000 + 1_000_000) * 128 + 1_000_000
        p = torch.zeros(P, device=dev); m = torch.zeros_like(p); v = torch.zeros_like(p)
        g = torch.full((P,), 1e-3, device=dev)
        k = [0]

        def adam():
            k[0] += 1
            _C.adamw_dense(p, m, v, g, 3e-4, 0.9, 0.999, 1e-8, 4e-5, k[0], zero_grad=True)
        adam()
        ms = timed(adam, 5)
        gbs = 32.0 * P / (ms * 1e-3) / 1e9   # 16 B read + 16 B written per parameter (g zeroed in the same pass)
        out['adamw_cfg4'] = {'params': P, 'ms': ms, 'bytes_per_param': 32, 'achieved_gbs': gbs, 'peak_gbs': peaks['hbm_gbs'],
                             'frac': gbs / peaks['hbm_gbs']}
        del p, m, v, g
    except Exception as ex:
        out['adamw_cfg4'] = {'error': repr(ex)}
    try:   # cfg4-shaped train step on one GPU: dense (torch-faithful) vs lazy (row-sparse) AdamW, reported separately
        from hassaku_b200.algorithms.sgd_alg import ArenaLayout
        from hassaku_b200.train.optim import DenseAdam
        U4, I4, d4, B4, N4 = 2_000_000, 1_000_000, 128, 8192, 50

        class _M:   # arena-only stand-in for the model (random-init weights of the cfg4 architecture, built on the device)
            pass
        mdl = _M()
        mdl.layout = ArenaLayout(U4, I4, d4, False, True, False)
        mdl.arena = torch.randn(mdl.layout.n_total, device=dev) * (0.1 / d4)
        mdl.parameters = lambda: [torch.nn.Parameter(mdl.arena[:4])]
        tabs = mdl.layout.tables(mdl.arena)
        gen = torch.Generator(device=dev); gen.manual_seed(0)
        ub = [torch.randint(0, U4, (B4,), device=dev, generator=gen) for _ in range(4)]
        ib = [torch.randint(0, I4, (B4, N4 + 1), device=dev, generator=gen) for _ in range(4)]
        acc = torch.zeros(1, dtype=torch.float64, device=dev)
        res4 = {}
        for mode in ('dense', 'lazy'):
            opt4 = DenseAdam(mdl, lr=3e-4, weight_decay=4e-5, mode=mode)
            kk = [0]

            def st():
                j = kk[0] % 4
                kk[0] += 1
                _C.mf_train_fused(tabs, opt4.grad_tables, ub[j], ib[j], 0, 0.0, acc)
                if mode == 'lazy':
                    opt4.mark(ub[j], ib[j])
                opt4.step_fused()
            st(); st()
            ms = timed(st, 6)
            ab4 = algorithmic_bytes(U4, I4, d4, B4, N4)
            res4[mode] = {'ms_per_step': ms, 'triples_per_s': B4 * N4 / (ms * 1e-3)}
            if mode == 'dense':
                res4[mode].update({'algorithmic_bytes': ab4['total'], 'achieved_gbs': ab4['total'] / (ms * 1e-3) / 1e9,
                                   'frac_of_hbm_peak': ab4['total'] / (ms * 1e-3) / 1e9 / peaks['hbm_gbs']})
            del opt4
        out['train_cfg4_1gpu'] = dict(res4, shape={'n_users': U4, 'n_items': I4, 'd': d4, 'B': B4, 'N': N4},
                                      note='lazy = row-sparse AdamW (different trajectory from torch.optim.AdamW), reported separately')
        del mdl, tabs
    except Exception as ex:
        out['train_cfg4_1gpu'] = {'error': repr(ex)}
    torch.cuda.empty_cache()
    try:
        import math
        from hassaku_b200.algorithms.sgd_alg import SGDMatrixFactorization
        from hassaku_b200.eval.eval import DeviceCSR, TopKScorer
        from scipy import sparse as sp
        U, I, d, B = 18944, 1_000_000, 256, 18944
        torch.manual_seed(0)
        model = SGDMatrixFactorization(U, I, d, use_item_bias=True)
        with torch.no_grad():
            for q in model.parameters():
                q.copy_(torch.randn_like(q) * (1.0 / math.sqrt(d) if q.shape[-1] == d else 0.05))
        model.to(dev)
        rng = np.random.RandomState(1)
        rows = np.repeat(np.arange(U), 80)
        ex = sp.csr_matrix((np.ones(len(rows), dtype=bool), (rows, rng.randint(0, I, len(rows)))), shape=(U, I))
        ex.sum_duplicates(); ex.sort_indices()
        ex = DeviceCSR(ex, dev)
        users = torch.arange(B, device=dev)
        fl = 2.0 * B * I * d
        for prec in ('bf16', 'tf32'):
            sc = TopKScorer(model, B, 100, prec)
            fn = lambda: sc(users, ex)
            fn(); fn()
            ms = timed(fn, 3)
            out[f'eval_tc_cfg5_{prec}'] = {'users': B, 'items': I, 'd': d, 'k': 100, 'ms_per_batch': ms, 'users_per_s': B / (ms * 1e-3),
                                           'tflops': fl / (ms * 1e-3) / 1e12, 'peak_tflops_sustained_bf16': peaks['bf16_tflops_sustained'],
                                           'frac_of_bf16_peak': fl / (ms * 1e-3) / 1e12 / peaks['bf16_tflops_sustained'],
                                           'includes': 'user-row pack + tcgen05 scoring + mask + top-100'}
            del sc
        del model
    except Exception as ex:
        out['eval_tc_cfg5'] = {'error': repr(ex)}
    torch.cuda.empty_cache()
    return out


def run_ours_sharded(args, wl):
    """Cleanup wrapper: captured CUDA graphs hold NCCL work and dist.destroy_process_group() blocks until they are
    dropped, so ShardedMF.close() must run even when the body raises (a stalled rank would hold the GPU box)."""
    import faulthandler
    import torch.distributed as dist
    # watchdog: a rank stalled in a collective dumps its Python stack to stderr and exits instead of holding the box
    faulthandler.dump_traceback_later(float(os.environ.get('HSK_BENCH_WATCHDOG', '900')), exit=True)
    holder = []
    try:
        _run_ours_sharded(args, wl, holder)
        faulthandler.cancel_dump_traceback_later()
    finally:
        for smf in holder:
            smf.close()
        if dist.is_initialized():
            dist.destroy_process_group()


def _run_ours_sharded(args, wl, holder):
    """N > 1: the item-/user-sharded step (hassaku_b200/sharded.py) with a per-GPU batch of `train_batch_size` samples
    (weak scaling: global batch = N x 8192), NCCL all-to-all for the row / gradient exchanges."""
    import torch
    import torch.distributed as dist
    from hassaku_b200.algorithms.sgd_alg import SGDMatrixFactorization
    from hassaku_b200.data.dataset import FullEvalDataset
    from hassaku_b200.data.synthetic import make_named
    from hassaku_b200.eval.eval import FullEvaluator
    from hassaku_b200.sharded import ShardedMF
    rank, world = int(os.environ['RANK']), int(os.environ['WORLD_SIZE'])
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local_rank)
    dev = torch.device('cuda', local_rank)
    dist.init_process_group('nccl', device_id=dev)
    name, d, B, N, loss, lr, wd = wl
    data = make_named(name)
    U, I = data.n_users, data.n_items
    K, W = args.steps, max(args.warmup, 0)
    torch.manual_seed(64)
    full = SGDMatrixFactorization(U, I, d, use_item_bias=True)
    smf = ShardedMF(U, I, d, use_item_bias=True, world=world, rank=rank, device=dev)
    holder.append(smf)
    smf.load_full_state_dict(full.state_dict())
    del full
    shift = float(np.log(I / N)) if loss == 'sampled_softmax' else 0.0
    # local batches: B samples whose user this rank owns
    import scipy.sparse as sp
    tr = data.train.tocsr()
    mine = sp.csr_matrix(tr[np.arange(rank, U, world)])
    sub = type('D', (), {})()
    sub.train, sub.n_items = sp.csr_matrix((mine.data, mine.indices, mine.indptr), shape=(mine.shape[0], I)), I
    us, its = make_batches(sub, B, N, 8, seed=64 + rank)
    us = [u * world + rank for u in us]                              # local row -> global user id
    u_dev = [torch.from_numpy(x).to(dev) for x in us]
    i_dev = [torch.from_numpy(x).to(dev) for x in its]
    Bg = B * world

    def barrier():
        dist.barrier()
        torch.cuda.synchronize()

    # the W warm-up steps asked for, and at least 30: the first replays of a freshly captured graph with NCCL nodes are slow
    for s in range(max(W, 30)):
        smf.step(u_dev[s % 8], i_dev[s % 8], Bg, loss, shift, lr, wd, exchange=args.exchange)
    smf.pop_loss()      # reset the loss accumulator (collective: every rank calls it)
    p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    # rank 0 samples the clocks of every GPU of the job; its start-up is over before the barrier that opens the region
    with ClockSampler(','.join(str(g) for g in range(world)), enabled=(rank == 0)) as clk:
        clk.wait_ready()
        barrier()
        # (i) pilot region: K steps, only to size the sustained run (reported as `pilot_ms_per_step`)
        p0.record()
        for s in range(K):
            smf.step(u_dev[s % 8], i_dev[s % 8], Bg, loss, shift, lr, wd, exchange=args.exchange)
        p1.record()
        barrier()
        # every step holds collectives, so step counts must be the SAME on every rank: derive them from max-over-ranks times
        t_pilot = torch.tensor([p0.elapsed_time(p1)], dtype=torch.float64, device=dev)
        dist.all_reduce(t_pilot, op=dist.ReduceOp.MAX)
        # (ii) sustained run (>= 1.5 s, untimed): clocks settle under load and the sampler sees it
        n_sus = max(K, int(1500.0 / max(float(t_pilot.item()) / K, 1e-3)))
        for s in range(n_sus):
            smf.step(u_dev[s % 8], i_dev[s % 8], Bg, loss, shift, lr, wd, exchange=args.exchange)
        barrier()
        # (iii) THE timed region: exactly K steps between barrier + synchronize, max over ranks
        a.record()
        for s in range(K):
            smf.step(u_dev[s % 8], i_dev[s % 8], Bg, loss, shift, lr, wd, exchange=args.exchange)
        b.record()
        barrier()
    t_all = torch.tensor([a.elapsed_time(b)], dtype=torch.float64, device=dev)
    dist.all_reduce(t_all, op=dist.ReduceOp.MAX)
    ms = float(t_all.item())
    last_loss = smf.pop_loss() / (2 * K + n_sus)
    # e2e: host batches
    u_pin = [torch.from_numpy(x).pin_memory() for x in us]
    i_pin = [torch.from_numpy(x).pin_memory() for x in its]
    a2, b2 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    a2.record()
    for s in range(K):
        smf.step(u_pin[s % 8].to(dev, non_blocking=True), i_pin[s % 8].to(dev, non_blocking=True), Bg, loss, shift, lr, wd,
                 exchange=args.exchange)
    loss_host = smf.pop_loss()
    b2.record()
    barrier()
    t2 = torch.tensor([a2.elapsed_time(b2)], dtype=torch.float64, device=dev)
    dist.all_reduce(t2, op=dist.ReduceOp.MAX)
    ms_e2e = float(t2.item())
    # sharded full-rank evaluation (first sweep warms NCCL's all-gather / all-to-all channels up)
    u2g = torch.from_numpy(data.user_group).float() if data.user_group is not None else None
    smf.evaluate(data.val, data.train, FullEvaluator(True, data.n_user_groups, u2g), batch_size=8192)
    barrier()
    ev_t0 = time.perf_counter()
    for _ in range(3):
        res = smf.evaluate(data.val, data.train, FullEvaluator(True, data.n_user_groups, u2g), batch_size=8192)
    torch.cuda.synchronize()
    eval_ms = (time.perf_counter() - ev_t0) * 1e3 / 3
    if rank == 0:
        ab = algorithmic_bytes(U, I, d, B, N)
        peak, peak_src = measured_peaks()
        triples = Bg * N
        cfg = workload_config(args.workload, wl, U, I)
        cfg.update({'parallelism': f'item+user row-sharded x{world}, NCCL ({args.exchange} exchange)', 'global_batch': Bg,
                    'l2': 'not flushed (back to back); tables are L2-resident'})
        line = {'metric': 'BPR-MF train triples/s', 'value': triples * K / (ms * 1e-3), 'unit': 'triples/s', 'n_gpus': world,
                'steps': K, 'warmup': W, 'ms_per_step': ms / K, 'higher_is_better': True, 'scaling': 'weak',
                'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic', 'config': cfg,
                'e2e': {'value': triples * K / (ms_e2e * 1e-3), 'unit': 'triples/s',
                        'h2d_bytes_per_step': int(us[0].nbytes + its[0].nbytes) * world, 'd2h_bytes_per_step': 8,
                        'ms_per_step': ms_e2e / K},
                'gpu_launches': 4 * K * world,
                'roofline': {'kernel': 'hsk_mf_train_fused_n', 'bound': 'hbm', 'achieved': None, 'peak': peak, 'unit': 'GB/s',
                             'frac': None, 'traffic': None, 'peak_source': peak_src,
                             'step': {'algorithmic_bytes_per_gpu': 2 * ab['A'] + 28 * ab['P'] // world,
                                      'note': 'per-kernel roofline is reported by the N = 1 run'}},
                'clocks': clk.summary(),
                'pilot_ms_per_step': float(t_pilot.item()) / K,
                'eval': {'metric': 'full-rank eval users/s', 'value': U / (eval_ms * 1e-3), 'unit': 'users/s',
                         'ms_per_sweep': eval_ms, 'ndcg@10': res['ndcg@10'], 'users': U},
                'final_loss': last_loss}
        print(json.dumps(line), flush=True)
    dist.barrier()


def run_ours(args, wl):
    import torch
    import torch.distributed as dist
    if int(os.environ.get('WORLD_SIZE', '1')) > 1:
        return run_ours_sharded(args, wl)
    from hassaku_b200 import _C
    from hassaku_b200.algorithms.sgd_alg import SGDMatrixFactorization
    from hassaku_b200.data.dataset import FullEvalDataset
    from hassaku_b200.data.synthetic import make_named
    from hassaku_b200.eval.eval import FullEvaluator, evaluate_recommender_algorithm
    from hassaku_b200.train.optim import DenseAdam
    from hassaku_b200.train.rec_losses import RecommenderSystemLossesEnum
    from hassaku_b200.train.trainer_step import FusedMFTrainStep

    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local_rank)
    dev = torch.device('cuda', local_rank)
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)
    name, d, B, N, loss, lr, wd = wl
    data = make_named(name)
    U, I = data.n_users, data.n_items
    K, W = args.steps, max(args.warmup, 0)

    init_like_reference(None)
    model = SGDMatrixFactorization(U, I, d, use_item_bias=True).to(dev)

    class _DS:
        n_items = I

    loss_fn = RecommenderSystemLossesEnum[loss].value.build_from_conf({'train_neg_strategy': 'uniform', 'neg_train': N},
                                                                      _DS())
    opt = DenseAdam(model, lr=lr, weight_decay=wd, decoupled=True)
    step = FusedMFTrainStep(model, loss_fn, opt)

    n_distinct = 16
    us, its = make_batches(data, B, N, n_distinct, seed=64 + rank)
    u_dev = [torch.from_numpy(x).to(dev) for x in us]
    i_dev = [torch.from_numpy(x).to(dev) for x in its]
    u_pin = [torch.from_numpy(x).pin_memory() for x in us]
    i_pin = [torch.from_numpy(x).pin_memory() for x in its]

    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)  # > 126 MB L2

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms):
        if world > 1:
            t = torch.tensor([ms], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            return float(t.item())
        return ms

    # ---- (1) resident inputs, L2 flushed between timed steps ----
    for s in range(W):
        step(u_dev[s % n_distinct], i_dev[s % n_distinct])
    ev0 = [torch.cuda.Event(enable_timing=True) for _ in range(K)]
    ev1 = [torch.cuda.Event(enable_timing=True) for _ in range(K)]
    barrier()
    clk = ClockSampler(local_rank)
    clk.__enter__()   # samples clocks / throttle reasons across every timed region below (stopped after region 2b)
    clk.wait_ready()
    barrier()
    for s in range(K):
        flush.zero_()
        ev0[s].record()
        step(u_dev[s % n_distinct], i_dev[s % n_distinct])
        ev1[s].record()
    barrier()
    ms_flushed = max_over_ranks(sum(a.elapsed_time(b) for a, b in zip(ev0, ev1)))

    # ---- (2) back to back (L2 warm), one event pair around K steps ----
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    a.record()
    for s in range(K):
        step(u_dev[s % n_distinct], i_dev[s % n_distinct])
    b.record()
    barrier()
    ms_warm = max_over_ranks(a.elapsed_time(b))

    # ---- (2b) sustained: back-to-back steps for >= 1.5 s so that the clock / power samples see the load ----
    n_sus = max(K, int(1500.0 / max(ms_warm / K, 1e-3)))
    a3, b3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a3.record()
    for s in range(n_sus):
        step(u_dev[s % n_distinct], i_dev[s % n_distinct])
    b3.record()
    barrier()
    ms_sus = max_over_ranks(a3.elapsed_time(b3))
    clk.__exit__()

    # ---- (3) per-kernel durations (events between the two launches), L2 flushed ----
    tabs, gtabs = model._tables(), opt.grad_tables
    kind, shift = _C.LOSS_KINDS[loss_fn.loss_kind], float(loss_fn.neg_shift())
    e = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(K)]
    barrier()
    for s in range(K):
        flush.zero_()
        e[s][0].record()
        _C.mf_train_fused(tabs, gtabs, u_dev[s % n_distinct], i_dev[s % n_distinct], kind, shift, step.loss_accum,
                          status=model._status())
        e[s][1].record()
        opt.step_fused()
        e[s][2].record()
    barrier()
    ms_fused = sum(x[0].elapsed_time(x[1]) for x in e) / K
    ms_adamw = sum(x[1].elapsed_time(x[2]) for x in e) / K

    # ---- (4) end to end through the public API with HOST batches ----
    loss_pin = torch.zeros(K, dtype=torch.float64).pin_memory()
    loss_dev = torch.zeros(K, dtype=torch.float64, device=dev)
    for s in range(min(W, 3)):
        step(u_pin[s % n_distinct], i_pin[s % n_distinct])
    a2, b2 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    t_wall0 = time.perf_counter()
    a2.record()
    for s in range(K):
        step(u_pin[s % n_distinct], i_pin[s % n_distinct], loss_out=loss_dev[s:s + 1])
        loss_pin[s:s + 1].copy_(loss_dev[s:s + 1], non_blocking=True)
    b2.record()
    barrier()
    wall_e2e = (time.perf_counter() - t_wall0) * 1e3
    ms_e2e = max_over_ranks(max(a2.elapsed_time(b2), 0.0))
    last_loss = float(loss_pin[K - 1])
    model.check_status()
    assert math.isfinite(last_loss), 'training diverged'

    # ---- (5) full-rank evaluation sweep (all users, top-100, 12 metrics) ----
    ds = FullEvalDataset.from_interactions(data.val, data.train, 'val', data.user_group, data.n_user_groups)

    class _Loader:
        dataset, batch_size = ds, 8192

    def eval_once():
        ev = FullEvaluator(True, ds.n_user_groups, ds.user_to_user_group)
        return evaluate_recommender_algorithm(model, _Loader, ev, dev)

    eval_once()
    barrier()
    t0 = time.perf_counter()
    n_eval = 5
    for _ in range(n_eval):
        res = eval_once()
    torch.cuda.synchronize()
    eval_ms = (time.perf_counter() - t0) * 1e3 / n_eval

    extras = {}
    if rank == 0 and not args.no_extras:
        extras = run_extras(dev, flush)
    if rank != 0:
        return
    ab = algorithmic_bytes(U, I, d, B, N)
    peak, peak_src = measured_peaks()
    # dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed ncu capture of this same command
    # (profiles/r01_ncu_step_kernels.md, L2 flushed before each launch): the tables are L2-resident, so DRAM traffic is far
    # BELOW the algorithmic bytes — the kernel is L2 / issue bound at this shape, not HBM bound
    ncu_traffic = {'hsk_mf_train_fused': 30.1e6, 'hsk_adamw_dense': 69.7e6}
    dom = 'hsk_mf_train_fused' if ms_fused >= ms_adamw else 'hsk_adamw_dense'
    dom_bytes = ab['gather_scatter'] if dom == 'hsk_mf_train_fused' else ab['adamw']
    dom_ms = max(ms_fused, ms_adamw)
    achieved = dom_bytes / (dom_ms * 1e-3) / 1e9
    triples = B * N * world
    line = {
        'metric': 'BPR-MF train triples/s', 'value': triples * K / (ms_flushed * 1e-3), 'unit': 'triples/s',
        'n_gpus': world, 'steps': K, 'warmup': W, 'ms_per_step': ms_flushed / K, 'higher_is_better': True,
        'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        'config': workload_config(args.workload, wl, U, I),
        'value_l2_warm': triples * K / (ms_warm * 1e-3), 'ms_per_step_l2_warm': ms_warm / K,
        'value_sustained': triples * n_sus / (ms_sus * 1e-3), 'sustained_steps': n_sus,
        'samples_per_s': B * world * K / (ms_flushed * 1e-3),
        'e2e': {'value': triples * K / (ms_e2e * 1e-3), 'unit': 'triples/s',
                'h2d_bytes_per_step': int(us[0].nbytes + its[0].nbytes), 'd2h_bytes_per_step': 8,
                'ms_per_step': ms_e2e / K, 'wall_ms_per_step': wall_e2e / K},
        'gpu_launches': 2 * K,
        'kernels_ms': {'hsk_mf_train_fused': ms_fused, 'hsk_adamw_dense': ms_adamw},
        'roofline': {'kernel': dom, 'bound': 'hbm', 'achieved': achieved, 'peak': peak, 'unit': 'GB/s',
                     'frac': achieved / peak, 'traffic': ncu_traffic[dom], 'peak_source': peak_src,
                     'algorithmic_bytes_per_launch': dom_bytes,
                     'step': {'algorithmic_bytes': ab['total'], 'achieved': ab['total'] / (ms_flushed / K * 1e-3) / 1e9,
                              'frac': ab['total'] / (ms_flushed / K * 1e-3) / 1e9 / peak},
                     'note': 'tables + optimizer state (63 MB) are L2-resident: algorithmic GB/s may exceed DRAM peak'},
        'clocks': clk.summary(),
        'eval': {'metric': 'full-rank eval users/s', 'value': U / (eval_ms * 1e-3), 'unit': 'users/s',
                 'ms_per_sweep': eval_ms, 'ndcg@10': res['ndcg@10'], 'users': U, 'timing': 'host wall clock incl. the '
                 'single D2H sync of the sweep'},
        'final_loss': last_loss,
        'extras': extras,
    }
    if not args.no_cpu_baseline:
        try:
            line['cpu_baseline'] = cpu_baseline(wl, data, us, its)
        except Exception as ex:  # never lose the GPU numbers to a baseline problem
            line['cpu_baseline'] = {'error': repr(ex)}
        try:
            line['eval']['cpu_baseline'] = cpu_eval_baseline(wl, data)
        except Exception as ex:
            line['eval']['cpu_baseline'] = {'error': repr(ex)}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=50)
    ap.add_argument('--warmup', type=int, default=5)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--workload', default='cfg2', choices=sorted(WORKLOADS))
    ap.add_argument('--no-cpu-baseline', action='store_true', help='skip the CPU leg (profiling runs)')
    ap.add_argument('--exchange', default='dense_graph', choices=['auto', 'dense', 'dense_graph', 'sparse'],
                    help='N > 1: item-row exchange of the sharded step')
    ap.add_argument('--no-extras', action='store_true', help='skip the cfg4 / cfg5-shaped kernel measurements')
    args = ap.parse_args()
    wl = WORKLOADS[args.workload]
    if args.impl == 'reference':
        run_reference(args, wl)
    else:
        run_ours(args, wl)


if __name__ == '__main__':
    main()
