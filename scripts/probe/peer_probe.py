"""NVLink read probe (measurement only): torchrun --nproc-per-node 2 scripts/probe/peer_probe.py
Every rank allocates a 1 GiB fp32 buffer of 512-byte rows, maps its neighbour's (hsk_peer_export / hsk_peer_open) and reads
LOCAL and PEER memory with the three access patterns of scripts/probe/peer_probe.cu, one rank at a time (unidirectional) and
all ranks at once (every link loaded in both directions).  Prints GB/s per GPU."""
import ctypes as C
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from hassaku_b200 import _C  # noqa: E402


def main():
    rank, world = int(os.environ['RANK']), int(os.environ['WORLD_SIZE'])
    torch.cuda.set_device(rank)
    dev = torch.device('cuda', rank)
    dist.init_process_group('nccl', device_id=dev)
    lib = C.CDLL(os.path.join(ROOT, 'gpurun_exp_probe', 'libpeer_probe.so'))
    n_rows_buf = 2 * 1024 * 1024                       # 1 GiB of 512-byte rows
    buf = torch.randn(n_rows_buf * 128, device=dev)
    out = torch.zeros(4, device=dev)
    exports = [None] * world
    dist.all_gather_object(exports, _C.peer_export(buf))
    nb = (rank + 1) % world
    peer = _C.peer_open(exports[nb][0], dev) + exports[nb][1]
    n_rows = 400_000                                    # ~ one step's rows (205 MB)
    gen = torch.Generator(device=dev); gen.manual_seed(rank)
    rows = torch.randint(0, n_rows_buf, (n_rows,), device=dev, generator=gen, dtype=torch.int32)
    st = torch.cuda.current_stream().cuda_stream
    sms = torch.cuda.get_device_properties(dev).multi_processor_count

    def timed(fn, n_bytes, both):
        for _ in range(2):
            fn()
        torch.cuda.synchronize(); dist.barrier()
        if not both and rank != 0:
            dist.barrier()
            return None
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(5):
            fn()
        b.record(); torch.cuda.synchronize()
        dist.barrier()
        return n_bytes * 5 / (a.elapsed_time(b) * 1e-3) / 1e9

    tests = []
    for name, ptr in (('local', buf.data_ptr()), ('peer', peer)):
        tests.append((f'{name} linear 1 GiB', lambda p=ptr: lib.probe_launch_linear(C.c_void_p(p), C.c_int64(buf.numel() * 4), sms * 8, C.c_void_p(out.data_ptr()), C.c_void_p(st)), buf.numel() * 4))
        for unroll in (2, 4, 8):
            for bps in (8, 16):
                tests.append((f'{name} rows ld   unroll {unroll} x {bps} CTAs/SM of 64 thr', lambda p=ptr, u=unroll, b=bps: lib.probe_launch_rows_ld(C.c_void_p(p), C.c_void_p(rows.data_ptr()), n_rows, u, sms * b, C.c_void_p(out.data_ptr()), C.c_void_p(st)), n_rows * 512))
        for stages in (4, 8, 16):
            for bps in (4, 8):
                tests.append((f'{name} rows bulk stages {stages} x {bps} CTAs/SM of 4 warps', lambda p=ptr, s_=stages, b=bps: lib.probe_launch_rows_bulk(C.c_void_p(p), C.c_void_p(rows.data_ptr()), n_rows, s_, sms * b, C.c_void_p(out.data_ptr()), C.c_void_p(st)), n_rows * 512))
    for name, fn, nbytes in tests:
        uni = timed(fn, nbytes, both=False)
        bi = timed(fn, nbytes, both=True)
        if rank == 0:
            print(f'{name:58s} one rank reading {uni:8.1f} GB/s | all ranks reading {bi:8.1f} GB/s per GPU', flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == '__main__':
    main()
