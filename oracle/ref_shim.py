"""TEST INFRASTRUCTURE ONLY — imports the *unmodified* reference (karapostK/hassaku) from
/root/reference so that golden vectors can be generated and the restatement in
`oracle/mf_oracle.py` can be validated against it.  Only usable in the build container:
/root/reference does not exist on the GPU box, and nothing in `-m gpu` tests, smoke() or
bench.py may import this module.

The three shims are import-time / version-drift fixes; no reference source is touched or copied:
  1. stub modules for `ray.air.session` (train/trainer.py:5), `matplotlib` (explanations/utils.py:5,11),
     `gdown` (data/data_utils.py:9), `wandb` is installed.
  2. `torch.utils.data.dataloader.T_co` alias (data/dataloader.py:8; renamed `_T_co` in torch>=2.x).
  3. scipy>=1.14 removed `.A` and rejects torch-tensor fancy indices (eval/eval.py:250) → adapter.
"""
import os
import sys
import types

REF_ROOT = os.environ.get("HASSAKU_REF_ROOT", "/root/reference")


def available() -> bool:
    return os.path.isdir(os.path.join(REF_ROOT, "algorithms"))


def _stub(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


class CsrCompat:
    """exclude_data[u_idxs.cpu()].A compatibility for scipy>=1.14 (eval/eval.py:250)."""

    def __init__(self, m):
        self.m = m

    def __getitem__(self, idx):
        import torch
        if isinstance(idx, torch.Tensor):
            idx = idx.cpu().numpy()
        return types.SimpleNamespace(A=self.m[idx].toarray())

    def __getattr__(self, item):
        return getattr(self.m, item)


_loaded = False


def load():
    """Make `algorithms`, `train`, `eval`, `data` (reference top-level packages) importable."""
    global _loaded
    if _loaded:
        return
    if not available():
        raise RuntimeError(f"reference tree not found at {REF_ROOT}")
    import torch.utils.data.dataloader as dl
    if not hasattr(dl, "T_co"):
        dl.T_co = dl._T_co
    if "ray" not in sys.modules:
        ray = _stub("ray")
        air = _stub("ray.air")
        ses = _stub("ray.air.session", report=lambda *a, **k: None)
        ray.air = air
        air.session = ses
    if "matplotlib" not in sys.modules:
        try:
            import matplotlib  # noqa: F401
        except Exception:
            mpl = _stub("matplotlib", use=lambda *a, **k: None)
            mpl.pyplot = _stub("matplotlib.pyplot")
    if "gdown" not in sys.modules:
        try:
            import gdown  # noqa: F401
        except Exception:
            _stub("gdown")
    os.environ.setdefault("WANDB_MODE", "disabled")
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
    import data.dataset as ds
    if not getattr(ds.FullEvalDataset, "_hsk_patched", False):
        _orig = ds.FullEvalDataset._prepare_data

        def _prep(self):
            _orig(self)
            self.exclude_data = CsrCompat(self.exclude_data)

        ds.FullEvalDataset._prepare_data = _prep
        ds.FullEvalDataset._hsk_patched = True
    _loaded = True
