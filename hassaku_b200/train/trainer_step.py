"""One fused training step of SGDMatrixFactorization (train/trainer.py:133-148 of the reference)."""
import torch

from hassaku_b200 import _C
from hassaku_b200.train.optim import DenseAdam
from hassaku_b200.train.rec_losses import RecommenderSystemLoss


class FusedMFTrainStep:
    """One training step of SGDMatrixFactorization: trainer.py:133-148 as hsk_mf_train_fused + hsk_adamw_dense."""

    def __init__(self, model, rec_loss: RecommenderSystemLoss, optimizer: DenseAdam):
        if rec_loss.loss_kind not in _C.LOSS_KINDS:
            raise ValueError(f'Loss {rec_loss.name} has no fused kernel')
        self.model, self.rec_loss, self.optimizer = model, rec_loss, optimizer
        self.kind = _C.LOSS_KINDS[rec_loss.loss_kind]
        self.shift = float(rec_loss.neg_shift())
        self.device = model.arena.device
        self.loss_accum = torch.zeros(1, dtype=torch.float64, device=self.device)

    def __call__(self, u_idxs: torch.Tensor, i_idxs: torch.Tensor, loss_out: torch.Tensor = None):
        """Enqueue one step.  `loss_out` (fp64 [1], device) receives this batch's mean loss added to it; by default
        the epoch accumulator is used."""
        u = u_idxs.to(self.device, torch.int64, non_blocking=True).contiguous()
        i = i_idxs.to(self.device, torch.int64, non_blocking=True).contiguous()
        _C.mf_train_fused(self.model._tables(), self.optimizer.grad_tables, u, i, self.kind, self.shift,
                          self.loss_accum if loss_out is None else loss_out, status=self.model._status())
        self.optimizer.step_fused()

    def pop_loss_sum(self) -> float:
        """Host sync: sum of the batch-mean losses since the last call (also surfaces index errors)."""
        v = float(self.loss_accum.item())
        self.loss_accum.zero_()
        self.model.check_status()
        return v
