"""sample_factory.model.model_utils (model/model_utils.py): building blocks user-defined torch modules import."""
from typing import List

from torch import nn


def get_rnn_size(cfg):
    """model_utils.py:11-24: placeholder 1 without an RNN, doubled for LSTM (h and c) and for separate actor / critic"""
    size = cfg.rnn_size * cfg.rnn_num_layers if cfg.use_rnn else 1
    if cfg.use_rnn and cfg.rnn_type == "lstm":
        size *= 2
    if not cfg.actor_critic_share_weights:
        size *= 2
    return size


def nonlinearity(cfg, inplace: bool = False) -> nn.Module:
    if cfg.nonlinearity == "elu":
        return nn.ELU(inplace=inplace)
    if cfg.nonlinearity == "relu":
        return nn.ReLU(inplace=inplace)
    if cfg.nonlinearity == "tanh":
        return nn.Tanh()
    raise Exception(f"Unknown {cfg.nonlinearity=}")


def fc_layer(in_features: int, out_features: int, bias=True, spec_norm=False) -> nn.Module:
    layer = nn.Linear(in_features, out_features, bias)
    return nn.utils.spectral_norm(layer) if spec_norm else layer


def create_mlp(layer_sizes: List[int], input_size: int, activation: nn.Module) -> nn.Module:
    layers = []
    for size in layer_sizes:
        layers.extend([fc_layer(input_size, size), activation])
        input_size = size
    return nn.Sequential(*layers) if layers else nn.Identity()


class ModelModule(nn.Module):
    def __init__(self, cfg):
        super().__init__()
        self.cfg = cfg

    def get_out_size(self):
        raise NotImplementedError()


def model_device(model):
    """device of a torch module's first parameter (model_utils.py)"""
    return next(model.parameters()).device
