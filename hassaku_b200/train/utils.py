"""Weight initialisation with the reference's distributions (train/utils.py:5-13 of karapostK/hassaku), so that the same
torch seed and module construction order give bit-identical initial weights:
embedding tables ~ N(0, (0.1 / width)^2) — the std suggested in the ProtoMF paper's appendix — and linear layers
Kaiming-uniform (relu gain) with zero bias.  Applied with `module.apply(general_weight_init)`; only exact
`nn.Embedding` / `nn.Linear` instances with trainable weights are touched."""
from torch import nn


def _init_embedding(table: nn.Embedding) -> None:
    width = table.weight.shape[-1]
    nn.init.normal_(table.weight, mean=0.0, std=0.1 / width)


def _init_linear(layer: nn.Linear) -> None:
    nn.init.kaiming_uniform_(layer.weight, nonlinearity='relu')
    bias = getattr(layer, 'bias', None)
    if bias is not None and bias.requires_grad:
        nn.init.constant_(bias, 0)


_INITIALISERS = {nn.Embedding: _init_embedding, nn.Linear: _init_linear}


def general_weight_init(m: nn.Module) -> None:
    init = _INITIALISERS.get(type(m))          # exact type match, like the reference (subclasses keep their own init)
    if init is not None and m.weight.requires_grad:
        init(m)
