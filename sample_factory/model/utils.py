"""sample_factory.model.utils (model/utils.py): weight-init helpers custom encoders apply to their torch modules"""
import torch.nn as nn


def _init(layer, weight_fn):
    if isinstance(layer, (nn.Linear, nn.Conv2d)):
        weight_fn(layer.weight)
        if layer.bias is not None:
            layer.bias.data.fill_(0)
    return layer


def orthogonal_init(layer, gain: float = 1.0):
    return _init(layer, lambda w: nn.init.orthogonal_(w, gain=gain))


def he_normal_init(layer):
    return _init(layer, lambda w: nn.init.kaiming_normal_(w, mode="fan_in", nonlinearity="relu"))
