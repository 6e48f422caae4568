#!/bin/bash
# One GPU call (1 x B200, ~4 min) that runs everything written after round 1's GPU budget was spent and the A/B
# measurements the opt-in kernels are waiting for.  Outputs go to gpurun_out/pending_*.
#   gpurun --timeout 600 -- 'bash scripts/validate_pending.sh'
# 2-GPU part (sharded tensor-core evaluation), separately:
#   gpurun --gpus 2 --timeout 300 -- 'HSK_RUN_UNVALIDATED=1 timeout 200 python -m pytest tests/test_gpu_sharded.py -q -m gpu'
# in-place exchange A/B (N = 2):
#   gpurun --gpus 2 --timeout 300 -- 'for v in 0 1; do HSK_SHARDED_INPLACE=$v HSK_BENCH_WATCHDOG=90 timeout 130 python -m torch.distributed.run \
#       --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 bench.py --gpus 2 --steps 200 --warmup 20 --no-extras --no-cpu-baseline; done'
set -u
mkdir -p gpurun_out
export HSK_RUN_UNVALIDATED=1
echo "== gated tests ==";  timeout 300 python -m pytest tests/test_gpu_next_rows.py -q -m gpu 2>&1 | tail -15 | tee gpurun_out/pending_tests.log
echo "== train ring kernel: shipping vs lean loop (cfg2) =="
timeout 120 python scripts/kbench.py train --workload cfg2 --variants tma,tma2 2>&1 | tail -4 | tee gpurun_out/pending_train_ab.log
echo "== tensor-core evaluator: shipping vs lean (cfg5-shaped batch) =="
for v in "" lean; do
  HSK_EVAL_TC=$v timeout 150 python scripts/kbench.py eval --users 18944 --items 1000000 --dim 256 --batch 18944 --variants bf16,tf32 --iters 5 2>&1 | tail -6 | sed "s/^/[HSK_EVAL_TC=$v] /"
done | tee gpurun_out/pending_eval_ab.log
echo "== cfg4-shaped sharded step at N = 1 (sparse exchange path on one rank) =="
timeout 120 python scripts/kbench_sharded.py --steps 20 2>&1 | tail -2 | tee gpurun_out/pending_sharded_cfg4.log
