#!/bin/bash
# Round-2 ncu evidence on one B200 (B200_PROFILING.md recipe): run under gpurun, read the .ncu-rep files back here with
# scripts/ncu_summary.py.  Numbers measured under ncu are never bench values.
#   gpurun --timeout 1200 -- 'bash scripts/profile_r02.sh'
set -u
mkdir -p gpurun_out
ARGS="--steps 20 --warmup 5 --no-also --no-cpu-baseline --eval-users 113664"
OURS='train_fused|adamw|mark_|sample_neg|eval_topk|pack_rows|rescore|rank_metrics|topk_'
echo "== 1. the command without a profiler (must exit 0) =="
timeout 300 python bench.py $ARGS > gpurun_out/r02_prof_plain.json 2> gpurun_out/r02_prof_plain.err || { tail -5 gpurun_out/r02_prof_plain.err; exit 1; }
echo "== 2. launch list of the same command (our kernels) =="
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"$OURS" -c 400 --csv \
    --log-file gpurun_out/r02_launches_bench_cfg4.csv python bench.py $ARGS > gpurun_out/r02_prof_launches.log 2>&1
echo "== 3. --set full: the train-step kernels (one launch each, 20 steps in) =="
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'train_fused|adamw_rows|mark_batch' --launch-skip 60 \
    --launch-count 3 -f -o gpurun_out/r02_ncu_step python bench.py --steps 20 --warmup 5 --no-also --no-cpu-baseline --no-eval \
    > gpurun_out/r02_prof_step.log 2>&1
echo "== 4. --set full: the evaluator kernels at the cfg5 shape (second batch) =="
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'eval_topk_tc2|rescore_topk|pack_rows' --launch-skip 4 \
    --launch-count 3 -f -o gpurun_out/r02_ncu_eval python bench.py --steps 5 --warmup 3 --no-also --no-cpu-baseline --eval-users 113664 \
    > gpurun_out/r02_prof_eval.log 2>&1
ls -la gpurun_out/r02_*
