"""Host-side mirror of the reference's algorithm ABCs (algorithms/base_classes.py:12-53, 88-165) — only what the
SGD-MF hot path needs.  Same names, argument meaning and error behaviour, so `Trainer`, `FullEvaluator` and
`experiment_helper.py`-style callers work unchanged."""
import logging
import os
from abc import ABC, abstractmethod
from typing import Dict, Tuple, Union

import torch
from torch import nn
from torch.utils import data


class RecommenderAlgorithm(ABC):
    """algorithms/base_classes.py:12-52."""

    def __init__(self):
        super().__init__()
        self.name = 'RecommenderAlgorithm'
        logging.info(f'Built {self.name} module')

    @abstractmethod
    def predict(self, u_idxs: torch.Tensor, i_idxs: torch.Tensor) -> torch.Tensor:
        """u_idxs [batch], i_idxs [batch, n_neg + 1] -> predictions [batch, n_neg + 1]."""

    @abstractmethod
    def save_model_to_path(self, path: str):
        pass

    @abstractmethod
    def load_model_from_path(self, path: str):
        pass

    @staticmethod
    @abstractmethod
    def build_from_conf(conf: dict, dataset: data.Dataset):
        pass


class SGDBasedRecommenderAlgorithm(RecommenderAlgorithm, ABC, nn.Module):
    """algorithms/base_classes.py:88-165."""

    def __init__(self):
        super().__init__()
        self.name = 'SGDBasedRecommenderAlgorithm'
        logging.info(f'Built {self.name} module')

    def forward(self, u_idxs: torch.Tensor, i_idxs: torch.Tensor) -> torch.Tensor:
        # base_classes.py:99-108
        u_repr = self.get_user_representations(u_idxs)
        i_repr = self.get_item_representations(i_idxs)
        return self.combine_user_item_representations(u_repr, i_repr)

    @abstractmethod
    def get_user_representations(self, u_idxs: torch.Tensor) -> Union[torch.Tensor, Tuple[torch.Tensor]]:
        pass

    @abstractmethod
    def get_item_representations(self, i_idxs: torch.Tensor) -> Union[torch.Tensor, Tuple[torch.Tensor, ...]]:
        pass

    @abstractmethod
    def combine_user_item_representations(self, u_repr, i_repr) -> torch.Tensor:
        pass

    def get_and_reset_other_loss(self) -> Dict:
        # base_classes.py:139-148
        return {'reg_loss': torch.zeros(1)}

    @torch.no_grad()
    def predict(self, u_idxs: torch.Tensor, i_idxs: torch.Tensor) -> torch.Tensor:
        # base_classes.py:150-154
        self.eval()
        return self(u_idxs, i_idxs)

    def save_model_to_path(self, path: str):
        # base_classes.py:156-159
        path = os.path.join(path, 'model.pth')
        torch.save(self.state_dict(), path)
        print('Model Saved')

    def load_model_from_path(self, path: str):
        # base_classes.py:161-165 (+ map_location so a GPU-trained model.pth loads anywhere)
        path = os.path.join(path, 'model.pth')
        state_dict = torch.load(path, map_location='cpu')
        self.load_state_dict(state_dict)
        print('Model Loaded')
