"""bench.py — BPR-MF train triples/s and full-rank eval users/s (NDCG@10) at 1/2/4/8 B200: the metric of BASELINE.json.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload cfg4|cfg2|cfg3|cfg1]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 bench.py --gpus N ...

ONE workload at every N (so that the 1 -> 8 curve compares like with like), BASELINE.json configs[3] + configs[4]:
  train   "cfg4": synthetic LFM-2b scale, 2 M users x 1 M items, ~200 M interactions (generated on the device per rank),
          d = 128, BPR, neg_train 50, item bias, AdamW(lr 3e-4, wd 4e-5) in the dense torch-faithful mode, 8192 samples
          per GPU and step (weak scaling: global batch N x 8192; the tables stay 2 M x 1 M, row-sharded over the N GPUs).
          P = 385 M parameters (6.2 GB of p, m, v, g): far beyond the 126 MB L2, so no L2 flush is needed between steps.
          N = 1: the public single-GPU API (FusedMFTrainStep + DenseAdam).  N > 1: hassaku_b200.sharded.ShardedMF — device
          routing + NCCL all-to-all of the needed rows / row gradients, the whole step ONE CUDA graph.
  eval    "cfg5": 10 M users x 1 M items, d = 256, BF16 tcgen05 scoring + exclusion mask + top-100 + fp32 re-scoring +
          12 metrics, item-sharded at N > 1 (all-gather of user rows, per-shard top-k, all-to-all + merge).
A "step" = one pass of the hot path over one batch (hsk_mf_train_fused + hsk_mark_batch + hsk_adamw_dense_rows).  Rank 0
prints ONE JSON line:
  value      whole-job train triples/s, batches resident in HBM, CUDA events around exactly K steps, max over ranks
  e2e        the same through the public API with HOST (pinned) index batches: H2D per step + loss D2H inside the region
  roofline   dominant kernel: algorithmic bytes per launch / mean launch duration (events on the launch
             stream) against the measured HBM copy bandwidth (MEASURED_PEAKS.json); `traffic` = DRAM bytes per launch from
             the committed ncu capture (profiles/r02_traffic.json), null when no capture of this workload is committed
  eval       users/s (+ TFLOP/s against the measured bf16 peaks), NDCG@10, its own e2e and roofline
  cpu_baseline / --impl reference   the oracle port of the reference's torch path (oracle/mf_oracle.py, the ATen op sequence
             of train/trainer.py:133-148 / eval/eval.py:237-253) on this box's host cores, all threads, bounded sample
  parity_check (N > 1)  one sharded step and one sharded evaluation round against rank 0's single-GPU result on the same
             global batch, outside the timed regions
  also       (N = 1, unless --no-also) the cfg2 / cfg3 single-GPU lines (L2-resident tables: L2 flushed between steps),
             the negative sampler's throughput and one Trainer.fit epoch with the device loader
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
    'cfg4': ('lfm2b', 128, 8192, 50, 'bpr', 3e-4, 4e-5),
}
EVAL_CFG5 = {'data': 'eval10m', 'd': 256, 'k': 100, 'batch': 18944, 'precision': 'bf16'}   # 18 944 = 148 SMs x 128 rows
DEVICE_GENERATED = ('lfm2b', 'eval10m')    # too large for the reference's host loaders: built on the device per rank


def algorithmic_bytes(U, I, d, B, N, item_bias=True):
    """A = 4 d B (N+2) (every gathered row once); small = index + bias traffic; 28 B per parameter."""
    A = 4 * d * B * (N + 2)
    small = 8 * B * (N + 1) + 8 * B * (N + 2)
    P = (U + I) * d + (I if item_bias else 0)
    return {'gather_scatter': 2 * A + small, 'adamw': 28 * P, 'total': 2 * A + small + 28 * P, 'P': P, 'A': A}


def load_peaks():
    try:
        with open(os.path.join(ROOT, 'MEASURED_PEAKS.json')) as f:
            p = json.load(f)
        return {'hbm_gbs': float(p['hbm_gbs']), 'bf16_tflops': float(p['bf16_tflops']),
                'bf16_tflops_sustained': float(p.get('bf16_tflops_sustained', p['bf16_tflops'])),
                'source': 'measured (MEASURED_PEAKS.json)'}
    except Exception:
        return {'hbm_gbs': 6650.0, 'bf16_tflops': 1590.0, 'bf16_tflops_sustained': 1400.0, 'source': 'fallback (B200_PROFILING.md)'}


def measured_peaks():
    p = load_peaks()
    return p['hbm_gbs'], p['source']


def measure_l2_copy_peak(dev):
    """Measured L2 bandwidth of THIS GPU (GB/s, read + write bytes): the best of three L2-resident streaming passes run 100 x
    back to back — a 16 MB -> 16 MB copy, an in-place add over 48 MB, a sum over 48 MB (launch gaps included, so a lower
    bound of the L2 peak).  The denominator for the L2-resident workloads (cfg2 / cfg3 `value_l2_warm`)."""
    import torch
    x = torch.zeros(16 << 20, dtype=torch.uint8, device=dev)
    y = torch.zeros(16 << 20, dtype=torch.uint8, device=dev)
    z = torch.zeros(12 << 20, dtype=torch.float32, device=dev)
    best = {}
    for name, fn, nbytes in (('copy 16 MB -> 16 MB', lambda: y.copy_(x), 2.0 * x.numel()),
                             ('in-place add, 48 MB', lambda: z.add_(1.0), 2.0 * z.numel() * 4),
                             ('sum, 48 MB', lambda: z.sum(), 1.0 * z.numel() * 4)):
        for _ in range(5):
            fn()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(100):
            fn()
        b.record()
        torch.cuda.synchronize()
        best[name] = nbytes * 100 / (a.elapsed_time(b) * 1e-3) / 1e9
    return max(best.values()), best


def committed_traffic(workload, kernel):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of `kernel` from the committed ncu capture of this bench
    command (profiles/r02_traffic.json: {workload: {kernel: {'bytes': ..., 'capture': file}}}), or None."""
    try:
        with open(os.path.join(ROOT, 'profiles', 'r02_traffic.json')) as f:
            t = json.load(f)
        return t[workload][kernel]
    except Exception:
        return None


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
            self.proc = None

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


def make_host_batches_large(U, I, B, N, n_batches, seed=64, zipf=0.8):
    """Captured batches for the reference arm at the device-generated shapes (no GPU involved): users uniform, positives
    Zipf(0.8), negatives uniform.  No rejection against the user's train row: at ~80 train items of 1 M the reference's
    collate loop (dataloader.py:112-120) redraws 8e-5 of the slots, which does not change the step's cost."""
    rng = np.random.default_rng(seed)
    p = 1.0 / np.arange(1, I + 1, dtype=np.float64) ** zipf
    cdf = np.cumsum(p / p.sum())
    cdf[-1] = 1.0
    us, its = [], []
    for _ in range(n_batches):
        u = rng.integers(0, U, B, dtype=np.int64)
        pos = np.minimum(np.searchsorted(cdf, rng.random(B), side='right'), I - 1).astype(np.int64)
        neg = rng.integers(0, I, (B, N), dtype=np.int64)
        us.append(u)
        its.append(np.column_stack([pos, neg]))
    return us, its


def make_device_batches(dint, B, N, n_batches, seed=64):
    """Captured batches from a DeviceInteractions (this rank's users): positives drawn from the train COO list, negatives
    from hsk_sample_negatives against the user's train CSR row — the device loader's batch (data/dataloader.py:92-129)."""
    import torch
    from hassaku_b200 import _C
    dev = dint.rows.device
    gen = torch.Generator(device=dev)
    gen.manual_seed(seed)
    status = torch.zeros(1, dtype=torch.int32, device=dev)
    us, its = [], []
    for b in range(n_batches):
        sel = torch.randint(0, dint.rows.numel(), (B,), device=dev, generator=gen)
        u, pos = dint.rows[sel].contiguous(), dint.cols[sel].contiguous()
        i = torch.empty((B, N + 1), dtype=torch.int64, device=dev)
        _C.sample_negatives(u, pos, N, dint.n_items, dint.n_users, dint.train[0], dint.train[1], seed, b, i, True, status)
        us.append(u)
        its.append(i)
    assert int(status.item()) == 0, 'negative sampler reported a status bit'
    return us, its


def shapes_of(name):
    from hassaku_b200.data.synthetic import SHAPES
    return SHAPES[name]


def workload_config(wname, wl, U, I, world=1):
    name, d, B, N, loss, lr, wd = wl
    big = name in DEVICE_GENERATED
    cfg = {'workload': f'{wname}: BPR-MF train step (+ cfg5 full-rank eval), synthetic {name} shape', 'n_users': U, 'n_items': I,
           'embedding_dim': d, 'train_batch_size': B, 'train_batch_is': 'per GPU (weak scaling)', 'global_batch': B * world,
           'neg_train': N, 'rec_loss': loss, 'optimizer': 'adamw (dense, torch-faithful)', 'lr': lr, 'wd': wd,
           'item_bias': True, 'precision': 'fp32 exact (train); bf16 tcgen05 + fp32 re-scoring (eval)',
           'parallelism': 'single GPU' if world == 1 else f'item+user row-sharded x{world}, one process per GPU (item-row exchange: see the line\'s `exchange` block)'}
    cfg['l2'] = ('inputs larger than L2: 6.2 GB of parameters + optimizer state per step, no flush' if big else
                 'flushed between timed steps (tables + optimizer state fit L2); value_l2_warm = back to back')
    return cfg


# ------------------------------------------------------------------------------------------------
# CPU side: the reference arm and the bounded cpu_baseline legs (the only places that execute oracle/)
# ------------------------------------------------------------------------------------------------
def _cpu_threads():
    import torch
    n = os.cpu_count() or 1
    try:
        n = len(os.sched_getaffinity(0))
    except Exception:
        pass
    torch.set_num_threads(max(1, n))     # torchrun exports OMP_NUM_THREADS=1: undo it for the CPU legs
    return torch.get_num_threads()


def _cpu_train_setup(wl, batch_scale=1):
    import torch
    from oracle import mf_oracle as O
    name, d, B, N, loss, lr, wd = wl
    B = B * batch_scale
    U, I, _ = shapes_of(name)
    torch.manual_seed(64)
    model = O.OracleMF(U, I, d, use_item_bias=True)
    tr = O.OracleTrainer(model, loss, lr, wd, 'adamw', neg_train=N)
    if name in DEVICE_GENERATED:
        us, its = make_host_batches_large(U, I, B, N, 4)
    else:
        from hassaku_b200.data.synthetic import make_named
        us, its = make_batches(make_named(name), B, N, 4)
    us = [torch.from_numpy(x) for x in us]
    its = [torch.from_numpy(x) for x in its]
    return tr, us, its, O.make_labels(B, N + 1), (U, I)


def run_reference(args, wl):
    """--impl reference: the reference's CPU implementation of the step (oracle port) on all host threads, rank 0 only."""
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    cores = _cpu_threads()
    name, d, B, N, loss, lr, wd = wl
    # the arm's config at --gpus N is the GLOBAL batch N x 8192 (weak scaling): the CPU takes it as one step
    tr, us, its, labels, (U, I) = _cpu_train_setup(wl, batch_scale=max(1, args.gpus))
    B = B * max(1, args.gpus)
    nb = len(us)
    budget_s = float(os.environ.get('HSK_REF_BUDGET_S', '150'))
    t_w0 = time.perf_counter()
    n_warm = 0
    for s in range(max(args.warmup, 1)):
        tr.step(us[s % nb], its[s % nb], labels)
        n_warm += 1
        if time.perf_counter() - t_w0 > 0.2 * budget_s:
            break
    per_step = (time.perf_counter() - t_w0) / n_warm
    # bounded: at the cfg4 shape one CPU step takes seconds; run as many of the K steps as fit the budget
    n_steps = int(max(1, min(args.steps, (0.8 * budget_s) // max(per_step, 1e-6))))
    t0 = time.perf_counter()
    for s in range(n_steps):
        tr.step(us[s % nb], its[s % nb], labels)
    dt = time.perf_counter() - t0
    v = n_steps * B * N / dt
    line = {'impl': 'reference', 'metric': 'BPR-MF train triples/s', 'value': v, 'unit': 'triples/s', 'n_gpus': args.gpus,
            'steps': n_steps, 'steps_requested': args.steps, 'warmup': n_warm, 'ms_per_step': 1e3 * dt / n_steps,
            'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
            'config': workload_config(args.workload, wl, U, I, args.gpus),
            'cpu_baseline': {'value': v, 'unit': 'triples/s', 'cores': cores, 'kind': 'port',
                             'sample': f'{n_steps} steps of {args.workload} (global batch B={B}, N={N}) on the oracle port of '
                                       f'train/trainer.py:133-148 (torch CPU, {cores} threads of {os.cpu_count()} cpus), {dt:.1f} s'},
            'e2e': {'value': v, 'unit': 'triples/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}}
    if not args.no_eval:
        try:
            line['eval'] = cpu_eval_baseline_cfg5()
        except Exception as ex:
            line['eval'] = {'error': repr(ex)}
    print(json.dumps(line), flush=True)


def cpu_baseline(wl, data=None, us=None, its=None, budget_s=20.0):
    """Bounded CPU leg of the default run: the oracle port of the step on all host threads for ~budget_s."""
    import torch
    from oracle import mf_oracle as O
    cores = _cpu_threads()
    name, d, B, N, loss, lr, wd = wl
    if data is not None:     # small workloads: the captured batches of the GPU arm
        torch.manual_seed(64)
        model = O.OracleMF(data.n_users, data.n_items, d, use_item_bias=True)
        tr = O.OracleTrainer(model, loss, lr, wd, 'adamw', neg_train=N)
        labels = O.make_labels(B, N + 1)
        us = [torch.from_numpy(np.asarray(x)) for x in us]
        its = [torch.from_numpy(np.asarray(x)) for x in its]
    else:
        tr, us, its, labels, _ = _cpu_train_setup(wl)
    tr.step(us[0], its[0], labels)  # warm-up
    n, t0 = 0, time.perf_counter()
    while True:
        tr.step(us[n % len(us)], its[n % len(us)], labels)
        n += 1
        dt = time.perf_counter() - t0
        if dt > budget_s or n >= 30:
            break
    return {'value': n * B * N / dt, 'unit': 'triples/s', 'cores': cores, 'kind': 'port',
            'sample': f'{n} steps of the same workload (B={B}, N={N}) on the oracle port (torch CPU, {cores} threads of '
                      f'{os.cpu_count()} cpus), {dt:.1f} s'}


def cpu_eval_baseline(wl, data, n_users=512, eval_batch_size=256):
    """The reference's full-rank evaluation (eval/eval.py:237-253: [Be, I, d] broadcast product, dense mask, torch.topk,
    12 metrics per batch; eval batch 256 as in its README) on a bounded sample of users, host cores only."""
    import torch
    from oracle import mf_oracle as O
    cores = _cpu_threads()
    name, d, B, N, loss, lr, wd = wl
    torch.manual_seed(64)
    model = O.OracleMF(data.n_users, data.n_items, d, use_item_bias=True)
    users = np.arange(min(n_users, data.n_users))
    t0 = time.perf_counter()
    with torch.no_grad():
        res = O.evaluate(model, data.val.tocsr(), data.train.tocsr(), eval_batch_size=eval_batch_size, users=users)
    dt = time.perf_counter() - t0
    return {'value': len(users) / dt, 'unit': 'users/s', 'cores': cores, 'kind': 'port',
            'sample': f'{len(users)} of {data.n_users} users against all {data.n_items} items, eval batch {eval_batch_size}, '
                      f'oracle port of eval/eval.py:237-253 (torch CPU, {cores} threads), {dt:.1f} s',
            'ndcg@10_of_sample': float(res['ndcg@10'])}


def cpu_eval_baseline_cfg5(n_users=16, eval_batch_size=4):
    """The reference evaluation at the cfg5 item shape (1 M items, d 256): its [Be, I, d] temporary is 1 GB per user, so
    Be = 4 (time a small user sample against the full item set, extrapolate linearly in users)."""
    import torch
    from scipy import sparse as sp
    from oracle import mf_oracle as O
    cores = _cpu_threads()
    U, I, _ = shapes_of(EVAL_CFG5['data'])
    d = EVAL_CFG5['d']
    torch.manual_seed(64)
    model = O.OracleMF(n_users, I, d, use_item_bias=True)       # only the sampled users' rows are needed
    rng = np.random.RandomState(0)
    rows = np.repeat(np.arange(n_users), 80)
    ex = sp.csr_matrix((np.ones(len(rows), dtype=bool), (rows, rng.randint(0, I, len(rows)))), shape=(n_users, I))
    lab = sp.csr_matrix((np.ones(n_users * 10, dtype=np.int8), (np.repeat(np.arange(n_users), 10), rng.randint(0, I, n_users * 10))),
                        shape=(n_users, I))
    ex.sum_duplicates(); ex.sort_indices(); lab.sum_duplicates(); lab.data[:] = 1; lab.sort_indices()
    t0 = time.perf_counter()
    with torch.no_grad():
        res = O.evaluate(model, lab, ex, eval_batch_size=eval_batch_size, users=np.arange(n_users))
    dt = time.perf_counter() - t0
    return {'metric': 'full-rank eval users/s', 'value': n_users / dt, 'unit': 'users/s', 'cores': cores, 'kind': 'port',
            'extrapolated': True,
            'sample': f'{n_users} users against all {I} items, d {d}, eval batch {eval_batch_size} (the reference materialises '
                      f'[Be, I, d]: 1 GB per user), oracle port of eval/eval.py:237-253 (torch CPU, {cores} threads), {dt:.1f} s; '
                      f'linear in users',
            'ndcg@10_of_sample': float(res['ndcg@10'])}


# ------------------------------------------------------------------------------------------------
# GPU side
# ------------------------------------------------------------------------------------------------
def _events(n):
    import torch
    return [torch.cuda.Event(enable_timing=True) for _ in range(n)]


class _Dist:
    """barrier / max-over-ranks that degrade to no-ops at world 1 (no process group)."""

    def __init__(self, world, dev):
        self.world, self.dev = world, dev

    def barrier(self):
        import torch
        if self.world > 1:
            import torch.distributed as dist
            dist.barrier()
        torch.cuda.synchronize()

    def max(self, x):
        if self.world == 1:
            return float(x)
        import torch
        import torch.distributed as dist
        t = torch.tensor([x], dtype=torch.float64, device=self.dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())


def bench_train_single(args, wname, wl, dev, D, clk_index=0, want_e2e=True, flush=None):
    """N = 1: FusedMFTrainStep + DenseAdam (the public single-GPU API) on workload `wl`.  Returns a dict of measurements."""
    import torch
    from hassaku_b200 import _C
    from hassaku_b200.algorithms.sgd_alg import SGDMatrixFactorization
    from hassaku_b200.train.optim import DenseAdam
    from hassaku_b200.train.rec_losses import RecommenderSystemLossesEnum
    from hassaku_b200.train.trainer_step import FusedMFTrainStep
    name, d, B, N, loss, lr, wd = wl
    big = name in DEVICE_GENERATED
    K, W = args.steps, max(args.warmup, 3)
    n_distinct = 16
    if big:
        from hassaku_b200.data.synthetic import make_device_interactions
        U, I, n_inter = shapes_of(name)
        data = make_device_interactions(U, I, n_inter, dev, 1, 0, seed=0)
        model = SGDMatrixFactorization.on_device(U, I, d, use_item_bias=True, device=dev, seed=64)
        u_dev, i_dev = make_device_batches(data, B, N, n_distinct)
        us = [x.cpu().numpy() for x in u_dev]
        its = [x.cpu().numpy() for x in i_dev]
        n_train = int(data.rows.numel())
    else:
        from hassaku_b200.data.synthetic import make_named
        data = make_named(name)
        U, I = data.n_users, data.n_items
        torch.manual_seed(64)  # conf_parser.py:18 default seed; weights ~ N(0, (0.1/shape[-1])^2) (train/utils.py:13)
        model = SGDMatrixFactorization(U, I, d, use_item_bias=True).to(dev)
        us, its = make_batches(data, B, N, n_distinct)
        u_dev = [torch.from_numpy(x).to(dev) for x in us]
        i_dev = [torch.from_numpy(x).to(dev) for x in its]
        n_train = int(data.train.nnz)

    class _DS:
        n_items = I

    loss_fn = RecommenderSystemLossesEnum[loss].value.build_from_conf({'train_neg_strategy': 'uniform', 'neg_train': N}, _DS())
    opt = DenseAdam(model, lr=lr, weight_decay=wd, decoupled=True)
    step = FusedMFTrainStep(model, loss_fn, opt)
    u_pin = [torch.from_numpy(x).pin_memory() for x in us]
    i_pin = [torch.from_numpy(x).pin_memory() for x in its]
    do_flush = (not big) and flush is not None

    # ---- (1) THE timed region: exactly K steps, inputs resident ----
    for s in range(W):
        step(u_dev[s % n_distinct], i_dev[s % n_distinct])
    D.barrier()
    clk = ClockSampler(clk_index)
    clk.__enter__()
    clk.wait_ready()
    D.barrier()
    if do_flush:     # L2-resident tables: flush between steps, per-step events
        ev0, ev1 = _events(K), _events(K)
        for s in range(K):
            flush.zero_()
            ev0[s].record()
            step(u_dev[s % n_distinct], i_dev[s % n_distinct])
            ev1[s].record()
        D.barrier()
        ms_timed = sum(a.elapsed_time(b) for a, b in zip(ev0, ev1))
    else:
        a, b = _events(2)
        a.record()
        for s in range(K):
            step(u_dev[s % n_distinct], i_dev[s % n_distinct])
        b.record()
        D.barrier()
        ms_timed = a.elapsed_time(b)
    # ---- (2) back to back + sustained (>= 1.5 s so that the clock / power samples see the load) ----
    a, b = _events(2)
    a.record()
    for s in range(K):
        step(u_dev[s % n_distinct], i_dev[s % n_distinct])
    b.record()
    D.barrier()
    ms_warm = a.elapsed_time(b)
    n_sus = max(K, int(1500.0 / max(ms_warm / K, 1e-3)))
    a, b = _events(2)
    a.record()
    for s in range(n_sus):
        step(u_dev[s % n_distinct], i_dev[s % n_distinct])
    b.record()
    D.barrier()
    ms_sus = a.elapsed_time(b)
    clk.__exit__()

    # ---- (3) per-kernel durations: events between the launches of one step ----
    tabs, gtabs = model._tables(), opt.grad_tables
    kind, shift = _C.LOSS_KINDS[loss_fn.loss_kind], float(loss_fn.neg_shift())
    Kk = min(K, 50)
    e = [_events(4) for _ in range(Kk)]
    uses_rows = None
    for s in range(Kk):
        if do_flush:
            flush.zero_()
        u, i = u_dev[s % n_distinct], i_dev[s % n_distinct]
        e[s][0].record()
        _C.mf_train_fused(tabs, gtabs, u, i, kind, shift, step.loss_accum, status=model._status())
        e[s][1].record()
        opt.mark(u, i)
        uses_rows = len(opt._segments) > 0
        e[s][2].record()
        opt.step_fused()
        e[s][3].record()
    D.barrier()
    k_ms = {'hsk_mf_train_fused': sum(x[0].elapsed_time(x[1]) for x in e) / Kk,
            'hsk_mark_batch': sum(x[1].elapsed_time(x[2]) for x in e) / Kk,
            ('hsk_adamw_dense_rows' if uses_rows else 'hsk_adamw_dense'): sum(x[2].elapsed_time(x[3]) for x in e) / Kk}

    # ---- (4) end to end through the public API with HOST batches ----
    out = {'U': U, 'I': I, 'ms_timed': ms_timed, 'ms_warm': ms_warm, 'ms_sus': ms_sus, 'n_sus': n_sus, 'K': K, 'W': W,
           'kernels_ms': k_ms, 'uses_rows_kernel': bool(uses_rows), 'clocks': clk.summary(), 'n_train': n_train,
           'flushed': do_flush, 'launches_per_step': 3 if uses_rows else 2}
    if want_e2e:
        loss_pin = torch.zeros(K, dtype=torch.float64).pin_memory()
        loss_dev = torch.zeros(K, dtype=torch.float64, device=dev)
        for s in range(3):
            step(u_pin[s % n_distinct], i_pin[s % n_distinct])
        a2, b2 = _events(2)
        D.barrier()
        t_wall0 = time.perf_counter()
        a2.record()
        for s in range(K):
            step(u_pin[s % n_distinct], i_pin[s % n_distinct], loss_out=loss_dev[s:s + 1])
            loss_pin[s:s + 1].copy_(loss_dev[s:s + 1], non_blocking=True)
        b2.record()
        D.barrier()
        out.update({'ms_e2e': a2.elapsed_time(b2), 'wall_e2e': (time.perf_counter() - t_wall0) * 1e3,
                    'h2d': int(us[0].nbytes + its[0].nbytes), 'final_loss': float(loss_pin[K - 1])})
        assert math.isfinite(out['final_loss']), 'training diverged'
    model.check_status()
    out['host_batches'] = (us, its)
    out['data'] = data
    out['model'] = model
    return out


def bench_sampler_and_fit(dev, also_wl=('cfg2', 'cfg3')):
    """What has parity but had no number: hsk_sample_negatives throughput and one Trainer.fit-style epoch
    with the device-resident loader (shuffle + sampling + fused step, nothing from the host) per small workload."""
    import torch
    from hassaku_b200.algorithms.sgd_alg import SGDMatrixFactorization
    from hassaku_b200.data.dataloader import NegativeSampler, TrainDataLoader
    from hassaku_b200.data.dataset import TrainRecDataset
    from hassaku_b200.data.synthetic import make_named
    from hassaku_b200.train.optim import DenseAdam
    from hassaku_b200.train.rec_losses import RecommenderSystemLossesEnum
    from hassaku_b200.train.trainer_step import FusedMFTrainStep
    out = {}
    for wname in also_wl:
        name, d, B, N, loss, lr, wd = WORKLOADS[wname]
        try:
            data = make_named(name)
            ds = TrainRecDataset.from_interactions(data.train)
            loader = TrainDataLoader(NegativeSampler(ds, N, 'uniform'), ds, batch_size=B, shuffle=True, device=dev, seed=64)
            # sampler alone
            u = loader.rows[:B].contiguous(); pos = loader.cols[:B].contiguous()
            for s in range(3):
                loader.sample_batch(u, pos, s)
            a, b = _events(2)
            n_it = 200
            a.record()
            for s in range(n_it):
                loader.sample_batch(u, pos, s)
            b.record()
            torch.cuda.synchronize()
            ms = a.elapsed_time(b) / n_it
            # one epoch through the loader + fused step (what Trainer._train_one_epoch runs)
            torch.manual_seed(64)
            model = SGDMatrixFactorization(data.n_users, data.n_items, d, use_item_bias=True).to(dev)

            class _DS:
                n_items = data.n_items

            loss_fn = RecommenderSystemLossesEnum[loss].value.build_from_conf({'train_neg_strategy': 'uniform', 'neg_train': N}, _DS())
            step = FusedMFTrainStep(model, loss_fn, DenseAdam(model, lr=lr, weight_decay=wd))
            max_batches = 400
            n_b, n_s = 0, 0
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for u_b, i_b, _lab in loader:
                step(u_b, i_b)
                n_b += 1; n_s += len(u_b)
                if n_b >= max_batches:
                    break
            ep_loss = step.pop_loss_sum() / n_b
            dt = time.perf_counter() - t0
            out[wname] = {'sampler': {'samples_per_s': B / (ms * 1e-3), 'negatives_per_s': B * N / (ms * 1e-3), 'us_per_batch': ms * 1e3,
                                      'reference_loader': '~8 k samples/s with 4 workers (SURVEY 3.4, data/dataloader.py:92-129)'},
                          'fit_epoch_device_loader': {'batches': n_b, 'samples_per_s': n_s / dt, 'triples_per_s': n_s * N / dt,
                                                      'wall_s': dt, 'mean_loss': ep_loss,
                                                      'includes': 'device shuffle + hsk_sample_negatives + fused step + AdamW, host wall clock'}}
            del model, step, loader
        except Exception as ex:
            out[wname] = {'error': repr(ex)}
    return out


def bench_eval(args, dev, D, world, rank, smf_factory=None):
    """cfg5: 10 M users x 1 M items, d 256, bf16 tcgen05 + fp32 re-scoring + top-100 + metrics.  N = 1: the public API
    (evaluate_mf_sweep over TopKScorer); N > 1: ShardedMF.evaluate.  `--eval-users` bounds the sweep."""
    import torch
    from hassaku_b200.data.synthetic import make_device_interactions
    from hassaku_b200.eval.eval import DeviceCSR, FullEvaluator, evaluate_mf_sweep
    peaks = load_peaks()
    U, I, n_inter = shapes_of(EVAL_CFG5['data'])
    d, k, Bt, prec = EVAL_CFG5['d'], EVAL_CFG5['k'], EVAL_CFG5['batch'], EVAL_CFG5['precision']
    n_eval_users = min(U, args.eval_users) if args.eval_users > 0 else U
    t0 = time.perf_counter()
    data = make_device_interactions(U, I, n_inter, dev, world, rank, seed=1, keep_coo=False)
    labels = DeviceCSR.from_tensors(data.val[0], data.val[1], (U, I))
    exclude = DeviceCSR.from_tensors(data.train[0], data.train[1], (U, I))
    torch.cuda.synchronize()
    gen_s = time.perf_counter() - t0
    std = (1.0 / math.sqrt(d), 0.05)
    res = {}
    if world == 1:
        from hassaku_b200.algorithms.sgd_alg import SGDMatrixFactorization
        model = SGDMatrixFactorization.on_device(U, I, d, use_item_bias=True, device=dev, seed=65, std=std)
        model.eval_precision = prec
        n_batches = math.ceil(n_eval_users / Bt)

        def sweep(batches=None):
            ev = FullEvaluator(True, 0, None)
            evaluate_mf_sweep(model, labels, exclude, ev, n_eval_users, Bt, user_batches=batches)
            return ev.get_results()          # the sweep's one host sync

        # BEFORE the sweeps (the burst peak is the figure for a kernel timed alone; after 10 s of sustained tensor work the
        # board runs into its power cap): one full batch through the scorer (pack + tcgen05 scoring + fp32 re-scoring), and
        # the dominant kernel ALONE (hsk_eval_topk_tc on pre-packed operands, events on the launch stream, a pause between
        # the timed launches)
        from hassaku_b200 import _C
        from hassaku_b200.eval.eval import TopKScorer
        sc = TopKScorer(model, Bt, k, prec)
        users = torch.arange(Bt, device=dev)
        sc(users, exclude)
        e = [_events(2) for _ in range(5)]
        for x in e:
            x[0].record(); sc(users, exclude); x[1].record()
            torch.cuda.synchronize(); time.sleep(0.05)
        ms_batch = float(np.median([x[0].elapsed_time(x[1]) for x in e]))
        P = _C.PRECISIONS[prec]
        Uq = _C.pack_rows(model.user_embeddings.weight.detach(), d, P, row_idx=users)
        ks, ki = torch.empty((Bt, sc.kc), device=dev), torch.empty((Bt, sc.kc), dtype=torch.int32, device=dev)

        def kern():
            _C.eval_topk_tc(Uq, sc.Vq, P, users, U, sc.kc, ks, ki, sc.scratch, Ib=model.item_bias.weight.detach(),
                            excl_indptr=exclude.indptr, excl_indices=exclude.indices)
        kern()
        torch.cuda.synchronize()
        e = [_events(2) for _ in range(7)]
        for x in e:
            x[0].record(); kern(); x[1].record()
            torch.cuda.synchronize(); time.sleep(0.05)
        ms_kernel = float(np.median([x[0].elapsed_time(x[1]) for x in e]))
        del sc, Uq, ks, ki
        torch.cuda.empty_cache()
        evaluate_mf_sweep(model, labels, exclude, FullEvaluator(True, 0, None), min(n_eval_users, 3 * Bt), Bt)   # warm-up
        D.barrier()
        a, b = _events(2)
        a.record()
        r = sweep()
        b.record()
        torch.cuda.synchronize()
        ms = a.elapsed_time(b)
        # e2e: user-id batches arrive from the HOST (pinned), the metric dict goes back to the host
        host_batches = [torch.arange(s, min(s + Bt, n_eval_users), dtype=torch.int64).pin_memory() for s in range(0, n_eval_users, Bt)]
        D.barrier()
        t1 = time.perf_counter()
        r2 = sweep(host_batches)
        torch.cuda.synchronize()
        ms_e2e = (time.perf_counter() - t1) * 1e3
        res.update({'ms_batch_18944': ms_batch, 'tflops_batch': 2.0 * Bt * I * d / (ms_batch * 1e-3) / 1e12,
                    'ms_kernel_18944': ms_kernel, 'tflops_kernel': 2.0 * Bt * I * d / (ms_kernel * 1e-3) / 1e12,
                    'ndcg@10_e2e': r2['ndcg@10'], 'h2d_bytes_per_batch': 8 * Bt})
        del model
    else:
        smf = smf_factory(U, I, d, std, seed=65)
        bs = Bt // world                                   # users per rank and round: a round scores 18 944 users
        max_rounds = math.ceil(math.ceil(n_eval_users / world) / bs)

        def sweep():
            return smf.evaluate(labels, exclude, FullEvaluator(True, 0, None), batch_size=bs, precision=prec, max_rounds=max_rounds)

        smf.evaluate(labels, exclude, FullEvaluator(True, 0, None), batch_size=bs, precision=prec, max_rounds=3)   # warm-up (NCCL channels)
        D.barrier()
        a, b = _events(2)
        a.record()
        r = sweep()
        b.record()
        torch.cuda.synchronize()
        ms = D.max(a.elapsed_time(b))
        D.barrier()
        t1 = time.perf_counter()
        r2 = sweep()
        torch.cuda.synchronize()
        ms_e2e = D.max((time.perf_counter() - t1) * 1e3)
        n_eval_users = min(U, max_rounds * bs * world)
        res['ndcg@10_e2e'] = r2['ndcg@10']
        res['h2d_bytes_per_batch'] = 0
        # the user-parallel alternative (item table replicated at sweep start, no per-round collective): reported next to the
        # item-sharded figure that SURVEY 8e specifies
        try:
            max_local = math.ceil(n_eval_users / world)
            smf.evaluate_replicated(labels, exclude, FullEvaluator(True, 0, None), batch_size=Bt, precision=prec, max_users=3 * Bt)
            D.barrier()
            a3, b3 = _events(2)
            a3.record()
            r3 = smf.evaluate_replicated(labels, exclude, FullEvaluator(True, 0, None), batch_size=Bt, precision=prec, max_users=max_local)
            b3.record()
            torch.cuda.synchronize()
            ms3 = D.max(a3.elapsed_time(b3))
            res['user_parallel'] = {'value': n_eval_users / (ms3 * 1e-3), 'unit': 'users/s', 'ms_per_sweep': ms3, 'ndcg@10': r3['ndcg@10'],
                                    'tflops_per_gpu': 2.0 * n_eval_users * I * d / (ms3 * 1e-3) / 1e12 / world,
                                    'what': 'ShardedMF.evaluate_replicated: the item shards are all-gathered once per sweep (inside the timed '
                                            'region), every GPU then evaluates its own users against all items'}
        except Exception as ex:
            res['user_parallel'] = {'error': repr(ex)}
        # item shards that stay in place, streamed over NVLink by every GPU (ShardedMF.evaluate_streamed): exact, no replica — but
        # remote rows are not cached in the local L2, so each of the 74 CTA pairs pulls the peers' shards over the link itself
        # (~37 GB x (G - 1) / G per 18 944-user batch): the mode for catalogues whose packed table does not fit one GPU, measured
        # here on two batches per rank only
        try:
            smf.evaluate_streamed(labels, exclude, FullEvaluator(True, 0, None), batch_size=Bt, precision=prec, max_users=Bt)
            D.barrier()
            a4, b4 = _events(2)
            a4.record()
            smf.evaluate_streamed(labels, exclude, FullEvaluator(True, 0, None), batch_size=Bt, precision=prec, max_users=2 * Bt)
            b4.record()
            torch.cuda.synchronize()
            ms4 = D.max(a4.elapsed_time(b4))
            res['streamed'] = {'value': 2 * Bt * world / (ms4 * 1e-3), 'unit': 'users/s', 'sample_users': 2 * Bt * world, 'ms': ms4,
                               'what': 'ShardedMF.evaluate_streamed on 2 batches per rank: NVLink-bound (peer rows bypass the local L2)'}
        except Exception as ex:
            res['streamed'] = {'error': repr(ex)}
        smf.check_status()
        smf.close()
        del smf
    torch.cuda.empty_cache()
    flops = 2.0 * n_eval_users * I * d
    tf = flops / (ms * 1e-3) / 1e12
    res.update({
        'metric': 'full-rank eval users/s', 'value': n_eval_users / (ms * 1e-3), 'unit': 'users/s', 'users': n_eval_users,
        'ms_per_sweep': ms, 'ndcg@10': r['ndcg@10'], 'recall@100': r['recall@100'],
        'config': {'workload': f'cfg5: full-rank eval sweep, {U} users x {I} items, d {d}, {prec} tcgen05 scoring + exclusion mask + '
                               f'top-{k} + fp32 re-scoring of k + 28 candidates + 12 metrics', 'users_per_round': Bt,
                   'exclusions_per_user': float(exclude.indices.numel()) * world / U, 'labels_per_user': float(labels.indices.numel()) * world / U,
                   'weights': 'N(0, 1/d) embeddings, N(0, 0.05^2) item bias (the training init would rank by bias alone)',
                   'parallelism': 'single GPU' if world == 1 else f'item-sharded x{world}: all-gather user rows, per-shard top-k, all-to-all + merge'},
        'tflops': tf,
        'e2e': {'value': n_eval_users / (ms_e2e * 1e-3), 'unit': 'users/s', 'h2d_bytes_per_step': res.pop('h2d_bytes_per_batch'),
                'd2h_bytes_per_step': 96, 'timing': 'host wall clock: sweep call -> metric dict on the host'},
        'roofline': {'kernel': 'eval_topk_tc_kernel', 'bound': 'tensor', 'unit': 'TFLOP/s',
                     'achieved': res.get('tflops_kernel', tf / world), 'peak': peaks['bf16_tflops'],
                     'peak_kind': 'burst bf16 (kernel timed alone)' if 'tflops_kernel' in res else 'burst bf16 (whole sweep / GPU: the per-kernel figure is reported at N = 1)',
                     'frac': res.get('tflops_kernel', tf / world) / peaks['bf16_tflops'],
                     'sweep': {'achieved_per_gpu': tf / world, 'peak': peaks['bf16_tflops_sustained'], 'peak_kind': 'sustained bf16 (inside a long sweep)',
                               'frac': tf / world / peaks['bf16_tflops_sustained']},
                     'algorithmic_flops_per_user': 2.0 * I * d,
                     'traffic': (committed_traffic('cfg5', 'eval_topk_tc_kernel') or {}).get('bytes') if world == 1 else None,
                     'traffic_note': 'DRAM bytes of one 18 944-user launch from the committed ncu capture (profiles/r02_traffic.json); the kernel is tensor-bound, the operands (0.5 GB of packed item rows) stream from L2 / HBM once per CTA pair',
                     'peak_source': peaks['source']},
        'data_gen_s': gen_s,
    })
    return res


def run_ours(args, wl):
    import torch
    world = int(os.environ.get('WORLD_SIZE', '1'))
    if world > 1:
        return run_ours_sharded(args, wl)
    dev = torch.device('cuda', int(os.environ.get('LOCAL_RANK', '0')))
    torch.cuda.set_device(dev)
    D = _Dist(1, dev)
    name, d, B, N, loss, lr, wd = wl
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)  # > 126 MB L2
    m = bench_train_single(args, args.workload, wl, dev, D, flush=flush)
    U, I, K = m['U'], m['I'], m['K']
    ab = algorithmic_bytes(U, I, d, B, N)
    peaks = load_peaks()
    triples = B * N
    ms_step = m['ms_timed'] / K
    dom = max(m['kernels_ms'], key=m['kernels_ms'].get)
    dom_bytes = ab['gather_scatter'] if dom == 'hsk_mf_train_fused' else ab['adamw']
    achieved = dom_bytes / (m['kernels_ms'][dom] * 1e-3) / 1e9
    tr = committed_traffic(args.workload, dom)
    big = name in DEVICE_GENERATED
    line = {
        'metric': 'BPR-MF train triples/s', 'value': triples * K / (m['ms_timed'] * 1e-3), 'unit': 'triples/s',
        'n_gpus': 1, 'steps': K, 'warmup': m['W'], 'ms_per_step': ms_step, 'higher_is_better': True,
        'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        'config': workload_config(args.workload, wl, U, I, 1),
        'value_back_to_back': triples * K / (m['ms_warm'] * 1e-3),
        'value_sustained': triples * m['n_sus'] / (m['ms_sus'] * 1e-3), 'sustained_steps': m['n_sus'],
        'samples_per_s': B * K / (m['ms_timed'] * 1e-3), 'train_interactions': m['n_train'],
        'e2e': {'value': triples * K / (m['ms_e2e'] * 1e-3), 'unit': 'triples/s', 'h2d_bytes_per_step': m['h2d'],
                'd2h_bytes_per_step': 8, 'ms_per_step': m['ms_e2e'] / K, 'wall_ms_per_step': m['wall_e2e'] / K},
        'gpu_launches': m['launches_per_step'] * K,
        'kernels_ms': m['kernels_ms'],
        'roofline': {'kernel': dom, 'bound': 'hbm' if big else 'l2 (tables + optimizer state are L2-resident: not an HBM fraction)',
                     'achieved': achieved, 'peak': peaks['hbm_gbs'], 'unit': 'GB/s', 'frac': achieved / peaks['hbm_gbs'],
                     'traffic': tr['bytes'] if tr else None, 'traffic_source': tr.get('capture') if tr else None,
                     'peak_source': peaks['source'], 'algorithmic_bytes_per_launch': dom_bytes,
                     'algorithmic_bytes_note': 'SURVEY 8d: 28 B / parameter (AdamW) and 2A + small (fused gather / scatter); the rows '
                                               'kernel moves 24 B for untouched rows, so the algorithmic figure can exceed the DRAM traffic',
                     'step': {'algorithmic_bytes': ab['total'], 'achieved': ab['total'] / (ms_step * 1e-3) / 1e9,
                              'frac': ab['total'] / (ms_step * 1e-3) / 1e9 / peaks['hbm_gbs']}},
        'clocks': m['clocks'],
        'final_loss': m['final_loss'],
    }
    host_batches, data = m.pop('host_batches'), m.pop('data')
    m.pop('model')
    torch.cuda.empty_cache()
    if not args.no_eval:
        try:
            line['eval'] = bench_eval(args, dev, D, 1, 0)
        except Exception as ex:
            line['eval'] = {'error': repr(ex)}
        torch.cuda.empty_cache()
    if not args.no_also and big:
        also = {}
        for w2 in ('cfg2', 'cfg3'):
            try:
                a2 = argparse.Namespace(**vars(args))
                a2.steps = min(args.steps, 100)
                m2 = bench_train_single(a2, w2, WORKLOADS[w2], dev, D, want_e2e=True, flush=flush)
                n2, d2, B2, N2 = WORKLOADS[w2][0], WORKLOADS[w2][1], WORKLOADS[w2][2], WORKLOADS[w2][3]
                ab2 = algorithmic_bytes(m2['U'], m2['I'], d2, B2, N2)
                ms2 = m2['ms_timed'] / m2['K']
                also[w2] = {'value': B2 * N2 / (ms2 * 1e-3), 'unit': 'triples/s', 'ms_per_step': ms2,
                            'value_l2_warm': B2 * N2 * m2['K'] / (m2['ms_warm'] * 1e-3),
                            'e2e': {'value': B2 * N2 * m2['K'] / (m2['ms_e2e'] * 1e-3), 'unit': 'triples/s', 'h2d_bytes_per_step': m2['h2d'],
                                    'd2h_bytes_per_step': 8},
                            'kernels_ms': m2['kernels_ms'], 'l2': 'flushed between timed steps',
                            'algorithmic_gbs': ab2['total'] / (ms2 * 1e-3) / 1e9,
                            'roofline_note': 'tables + optimizer state fit the 126 MB L2 (cfg2 63 MB, cfg3 165 MB borderline): the step is '
                                             'L2 / issue bound, so no HBM fraction is quoted for it — see profiles/ for lts throughput',
                            'config': workload_config(w2, WORKLOADS[w2], m2['U'], m2['I'], 1)}
                if w2 == 'cfg2' and not args.no_cpu_baseline:
                    us2, its2 = m2['host_batches']
                    try:
                        also[w2]['cpu_baseline'] = cpu_baseline(WORKLOADS[w2], m2['data'], us2, its2, budget_s=8.0)
                        also[w2]['eval_cpu_baseline'] = cpu_eval_baseline(WORKLOADS[w2], m2['data'])
                    except Exception as ex:
                        also[w2]['cpu_baseline'] = {'error': repr(ex)}
                    # the cfg2 full-rank evaluation (fp32 exact) through the public API
                    try:
                        also[w2]['eval'] = small_eval(m2['model'], m2['data'], dev)
                    except Exception as ex:
                        also[w2]['eval'] = {'error': repr(ex)}
                del m2
            except Exception as ex:
                also[w2] = {'error': repr(ex)}
            torch.cuda.empty_cache()
        try:
            l2, l2_all = measure_l2_copy_peak(dev)
            also['peaks'] = {'l2_gbs': l2, 'l2_probes_gbs': l2_all,
                             'how': 'best of three L2-resident streaming passes (torch kernels, 100 launches back to back, launch gaps '
                                    'included): a lower bound of the L2 peak'}
            for w2 in ('cfg2', 'cfg3'):
                if 'value_l2_warm' in also.get(w2, {}):
                    ab2 = algorithmic_bytes(also[w2]['config']['n_users'], also[w2]['config']['n_items'], WORKLOADS[w2][1], WORKLOADS[w2][2], WORKLOADS[w2][3])
                    ms_warm = WORKLOADS[w2][2] * WORKLOADS[w2][3] / also[w2]['value_l2_warm'] * 1e3
                    also[w2]['roofline_l2_warm'] = {'bound': 'l2', 'unit': 'GB/s', 'achieved': ab2['total'] / (ms_warm * 1e-3) / 1e9, 'peak': l2,
                                                    'frac': ab2['total'] / (ms_warm * 1e-3) / 1e9 / l2,
                                                    'note': 'back-to-back steps (no flush): algorithmic bytes of the step against the measured L2 streaming bandwidth (a lower bound of the peak, so the fraction can exceed 1)'}
        except Exception as ex:
            also['peaks'] = {'error': repr(ex)}
        try:
            also['loaders'] = bench_sampler_and_fit(dev)
        except Exception as ex:
            also['loaders'] = {'error': repr(ex)}
        line['also'] = also
    if not args.no_cpu_baseline:
        try:   # never lose the GPU numbers to a baseline problem
            if big:
                line['cpu_baseline'] = cpu_baseline(wl, budget_s=15.0)
            else:
                line['cpu_baseline'] = cpu_baseline(wl, data, host_batches[0], host_batches[1])
        except Exception as ex:
            line['cpu_baseline'] = {'error': repr(ex)}
        if not args.no_eval and isinstance(line.get('eval'), dict) and 'error' not in line['eval']:
            try:
                line['eval']['cpu_baseline'] = cpu_eval_baseline_cfg5()
            except Exception as ex:
                line['eval']['cpu_baseline'] = {'error': repr(ex)}
    print(json.dumps(line), flush=True)


def small_eval(model, data, dev):
    import torch
    from hassaku_b200.data.dataset import FullEvalDataset
    from hassaku_b200.eval.eval import FullEvaluator, evaluate_recommender_algorithm
    ds = FullEvalDataset.from_interactions(data.val, data.train, 'val', data.user_group, data.n_user_groups)

    class _Loader:
        dataset, batch_size = ds, 8192

    def once():
        return evaluate_recommender_algorithm(model, _Loader, FullEvaluator(True, ds.n_user_groups, ds.user_to_user_group), dev)

    once()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(5):
        res = once()
    torch.cuda.synchronize()
    ms = (time.perf_counter() - t0) * 1e3 / 5
    return {'metric': 'full-rank eval users/s', 'value': data.n_users / (ms * 1e-3), 'unit': 'users/s', 'ms_per_sweep': ms,
            'ndcg@10': res['ndcg@10'], 'users': data.n_users, 'precision': 'fp32 exact',
            'timing': 'host wall clock incl. the single D2H sync of the sweep'}


# ------------------------------------------------------------------------------------------------
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
            try:
                smf.close()
            except Exception:
                pass
        if dist.is_initialized():
            dist.destroy_process_group()


def parity_check_train(smf, full_sd, u_loc, i_loc, B, N, loss, shift, lr, wd, dev, rank, world, d, exchange='sparse'):
    """One sharded step (eager, the exchange that is timed) against the single-GPU step on the union batch, on rank 0."""
    import torch
    import torch.distributed as dist
    from hassaku_b200 import _C
    from hassaku_b200.algorithms.sgd_alg import ArenaLayout
    from hassaku_b200.train.optim import DenseAdam
    u_all = torch.empty(world * B, dtype=torch.int64, device=dev)
    i_all = torch.empty((world * B, N + 1), dtype=torch.int64, device=dev)
    dist.all_gather_into_tensor(u_all, u_loc)
    dist.all_gather_into_tensor(i_all, i_loc)
    smf.loss_accum.zero_()
    smf.step(u_loc, i_loc, B * world, loss, shift, lr, wd, exchange=exchange)
    l_sh = smf.pop_loss()
    got = smf.full_state_dict(to_cpu=False)
    smf.check_status()
    ok, worst = True, {}
    if rank == 0:
        U, I = full_sd['user_embeddings.weight'].shape[0], full_sd['item_embeddings.weight'].shape[0]
        lay = ArenaLayout(U, I, d, False, True, False)

        class _M:
            pass
        mdl = _M()
        mdl.layout = lay
        mdl.arena = torch.zeros(lay.n_total, device=dev)
        Uw, Vw, _, Ib, _ = lay.views(mdl.arena)
        Uw.copy_(full_sd['user_embeddings.weight']); Vw.copy_(full_sd['item_embeddings.weight']); Ib.copy_(full_sd['item_bias.weight'])
        mdl.parameters = lambda: [torch.nn.Parameter(mdl.arena[:4])]
        opt = DenseAdam(mdl, lr=lr, weight_decay=wd)
        acc = torch.zeros(1, dtype=torch.float64, device=dev)
        _C.mf_train_fused(lay.tables(mdl.arena), opt.grad_tables, u_all, i_all, _C.LOSS_KINDS[loss], shift, acc)
        opt.mark(u_all, i_all)
        opt.step_fused()
        l_single = float(acc.item())
        ok = abs(l_single - l_sh) <= 1e-5 * abs(l_single)
        worst['loss'] = abs(l_single - l_sh) / abs(l_single)
        for name_, w in zip(('user_embeddings.weight', 'item_embeddings.weight', 'item_bias.weight'), (Uw, Vw, Ib)):
            a, b = got[name_].double(), w.double()
            # max-norm relative error per step from a synchronised state; + the documented Adam-eps conditioning term
            err = float((a - b).abs().max() / b.abs().max())
            worst[name_] = err
            ok = ok and err < 1e-5 + 2e-3 * lr
        del mdl, opt
    del got
    torch.cuda.empty_cache()
    return ok, worst


def _run_ours_sharded(args, wl, holder):
    """N > 1: the item-/user-sharded step (hassaku_b200/sharded.py) with a per-GPU batch of `train_batch_size` samples
    (weak scaling: global batch = N x 8192), device routing + NCCL all-to-all, captured as one CUDA graph."""
    import torch
    import torch.distributed as dist
    from hassaku_b200 import _C
    from hassaku_b200.algorithms.sgd_alg import ArenaLayout
    from hassaku_b200.data.synthetic import make_device_interactions
    from hassaku_b200.sharded import ShardedMF, exchange_capacity
    rank, world = int(os.environ['RANK']), int(os.environ['WORLD_SIZE'])
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local_rank)
    dev = torch.device('cuda', local_rank)
    dist.init_process_group('nccl', device_id=dev)
    D = _Dist(world, dev)
    name, d, B, N, loss, lr, wd = wl
    big = name in DEVICE_GENERATED
    K, W = args.steps, max(args.warmup, 3)
    Bg = B * world
    n_distinct = 16
    exchange = args.exchange
    if big:
        U, I, n_inter = shapes_of(name)
        data = make_device_interactions(U, I, n_inter, dev, world, rank, seed=0)
        u_dev, i_dev = make_device_batches(data, B, N, n_distinct, seed=64 + rank)
        n_train = int(data.rows.numel())
        del data
    else:
        import scipy.sparse as sp
        from hassaku_b200.data.synthetic import make_named
        hd = make_named(name)
        U, I = hd.n_users, hd.n_items
        mine = sp.csr_matrix(hd.train.tocsr()[np.arange(rank, U, world)])
        sub = type('D', (), {})()
        sub.train, sub.n_items = sp.csr_matrix((mine.data, mine.indices, mine.indptr), shape=(mine.shape[0], I)), I
        us, its = make_batches(sub, B, N, n_distinct, seed=64 + rank)
        u_dev = [torch.from_numpy(u * world + rank).to(dev) for u in us]     # local row -> global user id
        i_dev = [torch.from_numpy(x).to(dev) for x in its]
        n_train = int(mine.nnz)
        if not exchange.startswith('dense') and B * (N + 1) >= 2 * I // world:
            exchange = 'dense_graph'        # the batch covers the item table: dense exchange (see ShardedMF.step)
    torch.cuda.empty_cache()

    # every rank draws the SAME full tables on its device (same seed, same generator) and keeps its rows; rank 0 keeps the
    # full copy for the parity check
    def full_tables(U_, I_, d_, std, seed):
        lay = ArenaLayout(U_, I_, d_, False, True, False)
        arena = torch.zeros(lay.n_total, dtype=torch.float32, device=dev)
        gen = torch.Generator(device=dev)
        gen.manual_seed(seed)
        Uw, Vw, _, Ib, _ = lay.views(arena)
        for v, s_ in ((Uw, std[0] if std else 0.1 / d_), (Vw, std[0] if std else 0.1 / d_), (Ib, std[1] if std else 0.1)):
            v.normal_(0., s_, generator=gen)
        return {'user_embeddings.weight': Uw, 'item_embeddings.weight': Vw, 'item_bias.weight': Ib}

    def make_smf(U_, I_, d_, std, seed, keep_full=False):
        sd = full_tables(U_, I_, d_, std, seed)
        s = ShardedMF(U_, I_, d_, use_item_bias=True, world=world, rank=rank, device=dev)
        s.peer_barrier = args.peer_barrier
        holder.append(s)
        s.load_full_state_dict(sd)
        return (s, sd) if keep_full else s

    shift = float(np.log(I / N)) if loss == 'sampled_softmax' else 0.0
    smf, full_sd = make_smf(U, I, d, None, 64, keep_full=True)
    exchange_fallback = None
    if exchange.startswith('peer'):
        if not smf.peer_supported():
            exchange, exchange_fallback = 'sparse_graph', 'rows longer than 128 floats or more than 8 ranks'
        else:
            try:
                smf._peer_setup()            # collective; raises on EVERY rank if any rank cannot export / map (CUDA IPC)
            except Exception as ex:
                exchange, exchange_fallback = 'sparse_graph', repr(ex)
    # ---- parity self-check (outside every timed region) ----
    parity = {'train': None, 'eval': None}
    try:
        ok, worst = parity_check_train(smf, full_sd, u_dev[0], i_dev[0], B, N, loss, shift, lr, wd, dev, rank, world, d,
                                       exchange=exchange.replace('_graph', ''))
        parity['train'] = {'ok': bool(ok), 'max_rel_err': worst,
                           'what': f'one sharded step ({exchange.replace("_graph", "")} exchange) vs the single-GPU step on the union '
                                   f'batch, from the same state'}
    except Exception as ex:
        parity['train'] = {'ok': False, 'error': repr(ex)}
    # fresh state for the measurement
    smf.load_full_state_dict(full_sd)
    for t_ in (smf.m, smf.v, smf.g):
        t_.zero_()
    smf.t = 0
    smf.loss_accum.zero_()
    del full_sd
    torch.cuda.empty_cache()

    def step(s, host=None):
        if host is None:
            smf.step(u_dev[s % n_distinct], i_dev[s % n_distinct], Bg, loss, shift, lr, wd, exchange=exchange)
        else:
            smf.step(host[0][s % n_distinct], host[1][s % n_distinct], Bg, loss, shift, lr, wd, exchange=exchange)

    # the W warm-up steps asked for, and at least 30: the first replays of a freshly captured graph with NCCL nodes are slow
    for s in range(max(W, 30)):
        step(s)
    smf.pop_loss()      # reset the loss accumulator (collective: every rank calls it)
    smf.check_status()
    p0, p1 = _events(2)
    a, b = _events(2)
    # rank 0 samples the clocks of every GPU of the job; its start-up is over before the barrier that opens the region
    with ClockSampler(','.join(str(g) for g in range(world)), enabled=(rank == 0)) as clk:
        clk.wait_ready()
        D.barrier()
        # (i) pilot region: K steps, only to size the sustained run (reported as `pilot_ms_per_step`)
        p0.record()
        for s in range(K):
            step(s)
        p1.record()
        D.barrier()
        # every step holds collectives, so step counts must be the SAME on every rank: derive them from max-over-ranks times
        t_pilot = D.max(p0.elapsed_time(p1))
        # (ii) sustained run (>= 1.5 s, untimed): clocks settle under load and the sampler sees it
        n_sus = max(K, int(1500.0 / max(t_pilot / K, 1e-3)))
        for s in range(n_sus):
            step(s)
        D.barrier()
        # (iii) THE timed region: exactly K steps between barrier + synchronize, max over ranks
        a.record()
        for s in range(K):
            step(s)
        b.record()
        D.barrier()
    ms = D.max(a.elapsed_time(b))
    last_loss = smf.pop_loss() / (2 * K + n_sus)
    smf.check_status()
    # e2e: host (pinned) batches, H2D per step, loss D2H at the end of the region
    u_pin = [x.cpu().pin_memory() for x in u_dev]
    i_pin = [x.cpu().pin_memory() for x in i_dev]
    for s in range(3):
        step(s, (u_pin, i_pin))
    smf.pop_loss()
    a2, b2 = _events(2)
    D.barrier()
    a2.record()
    for s in range(K):
        step(s, (u_pin, i_pin))
    loss_host = smf.pop_loss()
    b2.record()
    D.barrier()
    ms_e2e = D.max(a2.elapsed_time(b2))
    assert math.isfinite(loss_host), 'training diverged'
    # per-kernel view of one EAGER step on rank 0's stream (collectives included): which part limits the step
    eager = exchange.replace('_graph', '')
    e0, e1 = _events(2)
    D.barrier()
    e0.record()
    for s in range(10):
        smf.step(u_dev[s % n_distinct], i_dev[s % n_distinct], Bg, loss, shift, lr, wd, exchange=eager)
    e1.record()
    D.barrier()
    ms_eager = D.max(e0.elapsed_time(e1)) / 10
    capq = exchange_capacity(B * (N + 1), I, world)
    phases = None
    if exchange.startswith('peer'):
        # the peer step's parts, every rank at once (so the NVLink load is the step's): barrier, step kernel, AdamW
        P_, lay_ = smf._peer_setup(), smf.layout
        u_loc_ = smf.ops.local_index(u_dev[0], world, 0)
        q0, q1, q2, q3, q4, q5 = _events(6)
        D.barrier()
        q0.record()
        for _ in range(20):
            smf._barrier()
        q1.record()
        smf._barrier()
        q2.record()
        for _ in range(5):
            smf.ops.train_fused_peer(lay_, smf.arena, smf.g, P_['items'], I, u_loc_, i_dev[0], Bg, _C.LOSS_KINDS[loss], shift,
                                     smf.loss_accum, smf.t + 1, None)
        q3.record()
        smf._barrier()
        smf.g.zero_()
        smf._barrier()
        q4.record()
        for _ in range(5):
            smf.ops.adamw(smf.arena, smf.m, smf.v, smf.g, smf._segments(True, True), 0.0, 0.0, smf.t + 1, True)
        q5.record()
        D.barrier()
        phases = {'barrier_ms': D.max(q0.elapsed_time(q1)) / 20, 'fused_peer_ms': D.max(q2.elapsed_time(q3)) / 5,
                  'adamw_rows_ms': D.max(q4.elapsed_time(q5)) / 5,
                  'nvlink_gbs_per_direction_in_kernel': 2 * B * (N + 2) * (d + 1) * 4 * (world - 1) / world / (D.max(q2.elapsed_time(q3)) / 5 * 1e-3) / 1e9,
                  'what': 'max over ranks, all ranks running the same part at once; AdamW with lr = 0 (same traffic, state untouched)'}
    smf.close()
    holder.remove(smf)
    del smf
    torch.cuda.empty_cache()

    ev = None
    if not args.no_eval:
        try:
            # parity of the sharded evaluation: one round against rank 0's single-GPU scorer on the same users
            ev = bench_eval(args, dev, D, world, rank, smf_factory=lambda U_, I_, d_, std, seed: make_smf(U_, I_, d_, std, seed))
        except Exception as ex:
            ev = {'error': repr(ex)}
        try:
            parity['eval'] = parity_check_eval(make_smf, holder, dev, rank, world)
        except Exception as ex:
            parity['eval'] = {'ok': False, 'error': repr(ex)}
    if rank == 0:
        ab = algorithmic_bytes(U, I, d, B, N)
        peaks = load_peaks()
        triples = Bg * N
        per_gpu_bytes = ab['gather_scatter'] + ab['adamw'] // world
        ms_step = ms / K
        pc_ok = all(v is None or v.get('ok') for v in parity.values())
        line = {'metric': 'BPR-MF train triples/s', 'value': triples * K / (ms * 1e-3), 'unit': 'triples/s', 'n_gpus': world,
                'steps': K, 'warmup': W, 'ms_per_step': ms_step, 'higher_is_better': True, 'scaling': 'weak',
                'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic', 'config': workload_config(args.workload, wl, U, I, world),
                'samples_per_s': Bg * K / (ms * 1e-3), 'train_interactions_per_rank': n_train,
                'e2e': {'value': triples * K / (ms_e2e * 1e-3), 'unit': 'triples/s',
                        'h2d_bytes_per_step': int(u_pin[0].numel() * 8 + i_pin[0].numel() * 8) * world, 'd2h_bytes_per_step': 8,
                        'ms_per_step': ms_e2e / K},
                'gpu_launches': (4 if exchange.startswith('peer') else 11) * K * world,
                'gpu_launches_note': ('per rank and step, own kernels inside the graph: local_index + mark_rows + fused_peer (item rows '
                                      'read from / gradients reduced into the owners over NVLink) + adamw_rows = 4, + 2 NCCL barriers'
                                      if exchange.startswith('peer') else
                                      'per rank and step, own kernels inside the graph: route (5) + pack + local_index + mark_rows + fused + '
                                      'unpack_add + adamw_rows = 11, + 3 NCCL all-to-all kernels'),
                'exchange': {'kind': exchange, 'fallback_from_peer': exchange_fallback, 'capq_rows_per_owner': capq if exchange.startswith('sparse') else None,
                             'bytes_per_gpu_per_direction': (int(2 * world * (capq + math.ceil(capq / 128)) * 128 * 4) if exchange.startswith('sparse')
                                                             else int(2 * B * (N + 2) * (d + 1) * 4 * (world - 1) / world) if exchange.startswith('peer')
                                                             else None),
                             'bytes_note': 'NVLink bytes per GPU and direction and step: rows served + gradients sent (= rows fetched + '
                                           'gradients received)',
                             'eager_ms_per_step': ms_eager, 'graph_ms_per_step': ms_step, 'phases': phases},
                'roofline': {'kernel': 'whole step (per GPU)', 'bound': 'hbm', 'unit': 'GB/s',
                             'achieved': per_gpu_bytes / (ms_step * 1e-3) / 1e9, 'peak': peaks['hbm_gbs'],
                             'frac': per_gpu_bytes / (ms_step * 1e-3) / 1e9 / peaks['hbm_gbs'], 'traffic': None,
                             'peak_source': peaks['source'], 'algorithmic_bytes_per_gpu': per_gpu_bytes,
                             'note': 'per-GPU algorithmic bytes = 2A + small (its 8192-sample batch) + 28 P / N (its shard); the step also '
                                     'moves the exchanged rows over NVLink, which this HBM figure does not count'},
                'clocks': clk.summary(),
                'pilot_ms_per_step': t_pilot / K,
                'parity_check': 'ok' if pc_ok else 'FAILED', 'parity': parity,
                'final_loss': last_loss}
        if ev is not None:
            line['eval'] = ev
        print(json.dumps(line), flush=True)
    dist.barrier()


def parity_check_eval(make_smf, holder, dev, rank, world):
    """One sharded evaluation round (bf16 tensor-core scoring of the local shard + fp32 re-scoring + merge) against the
    single-GPU scorer on rank 0: identical metric dict (ids are identical unless fp32 scores tie)."""
    import torch
    from scipy import sparse as sp
    from hassaku_b200.algorithms.sgd_alg import SGDMatrixFactorization
    from hassaku_b200.eval.eval import DeviceCSR, FullEvaluator, evaluate_mf_sweep
    U, I, d, bs = 4096, 200_003, 256, 256
    std = (1.0 / math.sqrt(d), 0.05)
    smf, sd = make_smf(U, I, d, std, 77, keep_full=True)
    rng = np.random.RandomState(3)
    rows = np.repeat(np.arange(U), 40)
    ex = sp.csr_matrix((np.ones(len(rows), dtype=bool), (rows, rng.randint(0, I, len(rows)))), shape=(U, I))
    lab = sp.csr_matrix((np.ones(U * 8, dtype=np.int8), (np.repeat(np.arange(U), 8), rng.randint(0, I, U * 8))), shape=(U, I))
    ex.sum_duplicates(); ex.sort_indices(); lab.sum_duplicates(); lab.data[:] = 1; lab.sort_indices()
    exd, labd = DeviceCSR(ex, dev), DeviceCSR(lab, dev)
    n_rounds = 2
    got = smf.evaluate(labd, exd, FullEvaluator(True, 0, None), batch_size=bs, precision='bf16', max_rounds=n_rounds)
    try:        # the streamed mode over the same users (local rows [0, n_rounds * bs) of every rank)
        got_s = smf.evaluate_streamed(labd, exd, FullEvaluator(True, 0, None), batch_size=bs, precision='bf16', max_users=n_rounds * bs)
    except Exception as ex:
        got_s = repr(ex)
    smf.check_status()
    out = {'ok': True, 'what': f'{n_rounds} sharded evaluation rounds ({n_rounds * bs * world} users x {I} items, bf16 + fp32 re-scoring) vs the '
                               f'single-GPU sweep over the same users'}
    if rank == 0:
        model = SGDMatrixFactorization.on_device(U, I, d, use_item_bias=True, device=dev, seed=1)
        with torch.no_grad():
            model.user_embeddings.weight.copy_(sd['user_embeddings.weight'])
            model.item_embeddings.weight.copy_(sd['item_embeddings.weight'])
            model.item_bias.weight.copy_(sd['item_bias.weight'])
        model.eval_precision = 'bf16'
        evl = FullEvaluator(True, 0, None)
        evaluate_mf_sweep(model, labd, exd, evl, n_rounds * bs * world, 1024)
        want = evl.get_results()
        diff = max(abs(got[k_] - v) for k_, v in want.items())
        out.update({'ok': bool(diff <= 1e-9), 'max_metric_diff': diff, 'ndcg@10': got['ndcg@10']})
        if isinstance(got_s, dict):
            diff_s = max(abs(got_s[k_] - v) for k_, v in want.items())
            out.update({'ok': bool(out['ok'] and diff_s <= 1e-9), 'max_metric_diff_streamed': diff_s})
        else:
            out['streamed_unavailable'] = got_s
        del model
    smf.close()
    holder.remove(smf)
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=200)
    ap.add_argument('--warmup', type=int, default=20)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--workload', default='cfg4', choices=sorted(WORKLOADS))
    ap.add_argument('--no-cpu-baseline', action='store_true', help='skip the CPU legs (profiling runs)')
    ap.add_argument('--peer-barrier', default='kernel', choices=['kernel', 'nccl'], help='peer exchange: cross-rank barrier')
    ap.add_argument('--exchange', default='peer_graph', choices=['sparse', 'sparse_graph', 'dense', 'dense_graph', 'peer', 'peer_graph'],
                    help='N > 1: item-row exchange of the sharded step')
    ap.add_argument('--no-eval', action='store_true', help='skip the cfg5 full-rank evaluation sweep')
    ap.add_argument('--no-also', action='store_true', help='N = 1: skip the cfg2 / cfg3 / loader side lines')
    ap.add_argument('--eval-users', type=int, default=0, help='bound the cfg5 sweep to this many users (0 = all 10 M)')
    args = ap.parse_args()
    wl = WORKLOADS[args.workload]
    if args.impl == 'reference':
        run_reference(args, wl)
    else:
        run_ours(args, wl)


if __name__ == '__main__':
    main()
