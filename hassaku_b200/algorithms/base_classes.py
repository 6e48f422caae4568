"""Host-side mirror of the reference's algorithm ABCs (algorithms/base_classes.py:12-53, 88-165) — only what the
SGD-MF hot path needs.  Same names, argument meaning and error behaviour, so `Trainer`, `FullEvaluator` and
`experiment_helper.py`-style callers work unchanged."""
import logging
import os
from abc import ABC, abstractmethod
from typing import Dict, Tuple, Union

import torch
from torch import nn

MODEL_FILE = 'model.pth'      # checkpoint name inside `conf['model_path']` (base_classes.py:156-165)
Representation = Union[torch.Tensor, Tuple[torch.Tensor, ...]]


class RecommenderAlgorithm(ABC):
    """What every algorithm of the reference offers (algorithms/base_classes.py:12-52): scoring of (user, items) index
    batches, a checkpoint round trip through a directory, and a `build_from_conf(conf, dataset)` factory."""

    def __init__(self):
        super().__init__()
        self.name = 'RecommenderAlgorithm'
        logging.info('Built %s module', self.name)

    @abstractmethod
    def predict(self, u_idxs: torch.Tensor, i_idxs: torch.Tensor) -> torch.Tensor:
        """u_idxs [batch], i_idxs [batch, n_neg + 1] -> predictions [batch, n_neg + 1]."""

    @abstractmethod
    def save_model_to_path(self, path: str):
        """Write the model into directory `path`."""

    @abstractmethod
    def load_model_from_path(self, path: str):
        """Read back what `save_model_to_path` wrote."""

    @staticmethod
    @abstractmethod
    def build_from_conf(conf: dict, dataset):
        """Factory from the experiment configuration and the dataset (n_users, n_items, ...)."""


class SGDBasedRecommenderAlgorithm(RecommenderAlgorithm, ABC, nn.Module):
    """Algorithms trained by a `Trainer` with SGD (algorithms/base_classes.py:88-165): an `nn.Module` whose score is
    `combine(user representation, item representation)`."""

    def __init__(self):
        super().__init__()
        self.name = 'SGDBasedRecommenderAlgorithm'
        logging.info('Built %s module', self.name)

    # ---- the three pieces a model defines ----
    @abstractmethod
    def get_user_representations(self, u_idxs: torch.Tensor) -> Representation:
        """u_idxs [batch] -> the model's user representation(s)."""

    @abstractmethod
    def get_item_representations(self, i_idxs: torch.Tensor) -> Representation:
        """i_idxs [batch, n_neg + 1] -> the model's item representation(s)."""

    @abstractmethod
    def combine_user_item_representations(self, u_repr: Representation, i_repr: Representation) -> torch.Tensor:
        """-> logits [batch, n_neg + 1]."""

    # ---- defaults shared by all SGD models ----
    def forward(self, u_idxs: torch.Tensor, i_idxs: torch.Tensor) -> torch.Tensor:
        """Training-time scores (base_classes.py:99-108); SGDMatrixFactorization overrides it with the fused kernel."""
        return self.combine_user_item_representations(self.get_user_representations(u_idxs),
                                                      self.get_item_representations(i_idxs))

    def get_and_reset_other_loss(self) -> Dict:
        """Model-specific extra losses accumulated during `forward` (base_classes.py:139-148).  The Trainer adds
        `reg_loss` to the recommendation loss after every batch; plain MF has none."""
        return {'reg_loss': torch.zeros(1)}

    def predict(self, u_idxs: torch.Tensor, i_idxs: torch.Tensor) -> torch.Tensor:
        """Inference-time scores: eval mode, no autograd (base_classes.py:150-154)."""
        self.eval()
        with torch.no_grad():
            return self(u_idxs, i_idxs)

    def save_model_to_path(self, path: str):
        torch.save(self.state_dict(), os.path.join(path, MODEL_FILE))
        print('Model Saved')

    def load_model_from_path(self, path: str):
        # map_location: a model.pth written on a GPU loads on any host
        self.load_state_dict(torch.load(os.path.join(path, MODEL_FILE), map_location='cpu'))
        print('Model Loaded')
