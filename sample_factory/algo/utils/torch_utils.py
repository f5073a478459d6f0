"""sample_factory.algo.utils.torch_utils (algo/utils/torch_utils.py): helpers user model code calls."""
import torch


def calc_num_elements(module, module_input_shape):
    """number of output elements of `module` for one input of the given shape (torch_utils.py:36-40)"""
    return module(torch.rand((1,) + tuple(module_input_shape))).numel()


def to_scalar(value):
    return value.item() if isinstance(value, torch.Tensor) else value
