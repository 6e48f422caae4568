import os, sys, math, faulthandler
faulthandler.dump_traceback_later(40, exit=True)
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
os.environ.setdefault('MASTER_ADDR', '127.0.0.1'); os.environ.setdefault('MASTER_PORT', '29533')
rank = int(os.environ.get('RANK', 0)); world = int(os.environ.get('WORLD_SIZE', 1))
torch.cuda.set_device(rank)
dist.init_process_group('nccl', rank=rank, world_size=world, device_id=torch.device('cuda', rank))
from hassaku_b200.algorithms.sgd_alg import SGDMatrixFactorization
from hassaku_b200.sharded import ShardedMF
U, I, d, B, N = 801, 507, 128, 256, 10
dev = torch.device('cuda', rank)
torch.manual_seed(5)
full = SGDMatrixFactorization(U, I, d, use_item_bias=True)
b = ShardedMF(U, I, d, use_item_bias=True, world=world, rank=rank, device=dev)
b.load_full_state_dict(full.state_dict())
rng = np.random.RandomState(rank)
for s in range(3):
    u = torch.from_numpy((rng.randint(0, (U - rank + world - 1) // world, B) * world + rank).astype(np.int64)).to(dev)
    i = torch.from_numpy(rng.randint(0, I, (B, N + 1)).astype(np.int64)).to(dev)
    print('step', s, flush=True)
    b.step(u, i, B * world, 'bpr', 0.0, 1e-3, 1e-4, exchange='dense_graph')
    torch.cuda.synchronize()
    print('done', s, b.pop_loss(), flush=True)
b.close()
dist.destroy_process_group()
print('OK', flush=True)
