"""2-rank debug harness for the sharded step: every rank logs progress markers and dumps its Python stack if it stalls.
    python scripts/debug_sharded.py [world] [exchange]"""
import faulthandler
import math
import os
import sys
import time

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def worker(rank, world, exchange):
    os.makedirs('gpurun_out', exist_ok=True)
    log = open(f'gpurun_out/dbg_rank{rank}.txt', 'w')
    faulthandler.dump_traceback_later(45, exit=True, file=log)

    def mark(msg):
        log.write(f'{time.time():.3f} {msg}\n'); log.flush()
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT='29533')
    torch.cuda.set_device(rank)
    dist.init_process_group('nccl', rank=rank, world_size=world, device_id=torch.device('cuda', rank))
    mark('init done')
    dist.barrier(); torch.cuda.synchronize(); mark('barrier done')
    from hassaku_b200.sharded import ShardedMF, partition_batch_by_user_owner
    U, I, d, B, N = 6040, 3706, 402, 8192, 50
    dev = torch.device('cuda', rank)
    smf = ShardedMF(U, I, d, use_item_bias=True, world=world, rank=rank, device=dev)
    mark('model built')
    g = torch.Generator(device=dev); g.manual_seed(1)
    u = torch.randint(0, U, (B * world,), device=dev, generator=g)
    i = torch.randint(0, I, (B * world, N + 1), device=dev, generator=g)
    ul, il = partition_batch_by_user_owner(u, i, world, rank)
    torch.cuda.synchronize(); mark(f'partitioned {tuple(il.shape)}')
    for s in range(5):
        smf.step(ul[:B // 2], il[:B // 2], B, 'bpr', 0.0, 1e-3, 1e-4, exchange=exchange)
        torch.cuda.synchronize(); mark(f'step {s} done')
    mark(f'loss {smf.pop_loss()}')
    smf.close()
    dist.destroy_process_group()
    mark('closed')


if __name__ == '__main__':
    world = int(sys.argv[1]) if len(sys.argv) > 1 else 2
    exchange = sys.argv[2] if len(sys.argv) > 2 else 'dense'
    mp.spawn(worker, args=(world, exchange), nprocs=world, join=True)
