"""Opt-in NVTX ranges around the hot path's host-side phases (tracing), for nsys / ncu --nvtx timelines.

    import hassaku_b200.nvtx as hnvtx; hnvtx.enable()        # or HSK_NVTX=1 in the environment

Ranges: `hsk.train_step` (FusedMFTrainStep), `hsk.sharded_step` / `hsk.sharded_eval` / `hsk.sharded_eval_round`
(ShardedMF), `hsk.eval_sweep` / `hsk.eval_batch` (evaluate_mf_sweep).  Disabled (the default) a range is one attribute
test; enabled it is torch.cuda.nvtx.range_push / range_pop (about a microsecond each)."""
import contextlib
import os

import torch

_enabled = os.environ.get('HSK_NVTX') == '1'


def enable(on: bool = True):
    global _enabled
    _enabled = bool(on)


def enabled() -> bool:
    return _enabled


class _Null:
    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False


_NULL = _Null()


def range(name: str):   # noqa: A001 - mirrors torch.cuda.nvtx.range
    if not _enabled:
        return _NULL
    return torch.cuda.nvtx.range(name)
