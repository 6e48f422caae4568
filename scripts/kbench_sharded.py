"""cfg4-shaped sharded train step at N GPUs (BASELINE.json configs[3]: 2 M users x 1 M items, d 128, BPR, item table
row-sharded, NCCL all-to-all): the configuration the north star's "near-linear item-sharded scaling" is quoted on.
bench.py's contract line stays cfg2; this script measures the sparse-exchange step where the tables do not fit L2 and the
dense AdamW (28 B / parameter / step, sharded 1 / N) dominates.  Random-init tables, uniform random indices generated on
the device (no 200 M-interaction dataset is built).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 scripts/kbench_sharded.py \
        [--users 2000000 --items 1000000 --dim 128 --batch 8192 --neg 50 --steps 50 --exchange sparse]

Prints one JSON line on rank 0: ms/step (CUDA events, max over ranks), triples/s of the whole job (weak scaling: `batch`
samples per GPU), bytes of AdamW state streamed per GPU.  NOT YET RUN on GPUs (written after round 1's GPU budget)."""
import argparse
import faulthandler
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--users', type=int, default=2_000_000)
    ap.add_argument('--items', type=int, default=1_000_000)
    ap.add_argument('--dim', type=int, default=128)
    ap.add_argument('--batch', type=int, default=8192, help='samples per GPU')
    ap.add_argument('--neg', type=int, default=50)
    ap.add_argument('--steps', type=int, default=50)
    ap.add_argument('--warmup', type=int, default=5)
    ap.add_argument('--exchange', default='sparse', choices=['sparse', 'dense', 'auto'])
    args = ap.parse_args()
    faulthandler.dump_traceback_later(300, exit=True)     # a stalled collective must not hold the box
    rank, world = int(os.environ.get('RANK', 0)), int(os.environ.get('WORLD_SIZE', 1))
    local = int(os.environ.get('LOCAL_RANK', rank))
    os.environ.setdefault('MASTER_ADDR', '127.0.0.1'); os.environ.setdefault('MASTER_PORT', '29541')
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    dist.init_process_group('nccl', rank=rank, world_size=world, device_id=dev)
    from hassaku_b200.sharded import ShardedMF
    smf = None
    try:
        U, I, d, B, N = args.users, args.items, args.dim, args.batch, args.neg
        smf = ShardedMF(U, I, d, use_item_bias=True, world=world, rank=rank, device=dev)
        gen = torch.Generator(device=dev); gen.manual_seed(64 + rank)
        smf.arena.copy_(torch.randn(smf.arena.shape, device=dev, generator=gen) * 0.05)
        n_local_users = smf.spec.n_local_users
        us = [torch.randint(0, n_local_users, (B,), device=dev, generator=gen) * world + rank for _ in range(4)]
        its = [torch.randint(0, I, (B, N + 1), device=dev, generator=gen) for _ in range(4)]
        Bg = B * world
        for s in range(args.warmup):
            smf.step(us[s % 4], its[s % 4], Bg, 'bpr', 0.0, 3e-4, 4e-5, exchange=args.exchange)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        dist.barrier(); torch.cuda.synchronize()
        a.record()
        for s in range(args.steps):
            smf.step(us[s % 4], its[s % 4], Bg, 'bpr', 0.0, 3e-4, 4e-5, exchange=args.exchange)
        b.record()
        dist.barrier(); torch.cuda.synchronize()
        t = torch.tensor([a.elapsed_time(b)], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item()) / args.steps
        loss = smf.pop_loss() / (args.steps + args.warmup)
        assert int(smf.status.item()) == 0
        if rank == 0:
            print(json.dumps({'workload': 'cfg4-shaped sharded step', 'n_gpus': world, 'users': U, 'items': I, 'd': d,
                              'batch_per_gpu': B, 'neg': N, 'exchange': args.exchange, 'ms_per_step': ms,
                              'triples_per_s': Bg * N / (ms * 1e-3), 'adamw_bytes_per_gpu': 28 * smf.layout.n_total,
                              'mean_loss': loss}), flush=True)
    finally:
        if smf is not None:
            smf.close()
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
