"""oracle/ — TEST INFRASTRUCTURE, NOT PRODUCT.

A CPU restatement of the karapostK/hassaku SGD-MF hot path (train step + full-rank evaluator),
used ONLY as the checker by `tests/`, `__graft_entry__.smoke()` and the `cpu_baseline` /
`--impl reference` legs of `bench.py`.  Nothing under `hassaku_b200/` imports it; the product
path fails loudly if the CUDA library is missing instead of falling back to this code.

Parity status: PINNED.
  * `eval/metrics.py` — pinned by the reference's own 15 known answers
    (framework_tests/eval/test_metrics.py:29-69) -> tests/test_oracle_golden.py.
  * everything else (forward, losses, backward, AdamW, masking, top-k, FullEvaluator aggregation)
    is not covered by any reference test, so it is pinned by running the *unmodified* reference
    (imported from /root/reference through `oracle/ref_shim.py`) on seeded inputs and committing
    its outputs as fixtures: `oracle/make_golden.py` -> tests/golden/*.npz.

Files:
  mf_oracle.py   torch-CPU restatement (same ATen op sequence as the reference's Python)
  philox.py      numpy restatement of Philox4x32-10 and of the device negative sampler's index
                 spec (integer arithmetic, bit-exact contract), pinned by Random123 known answers
                 (the AdamW arithmetic has no C restatement here: its checker is torch.optim.AdamW /
                 Adam / Adagrad themselves, bitwise, in tests/test_gpu_adamw.py)
  ref_shim.py    import shims for the real reference (build container only)
  make_golden.py generates tests/golden/ from the real reference (build container only)
"""
