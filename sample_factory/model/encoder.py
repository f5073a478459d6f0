"""sample_factory.model.encoder (model/encoder.py:15-31): the base class custom encoders derive from.

The built-in encoders (MlpEncoder, ConvEncoder = convnet_simple / convnet_impala / convnet_atari) run as libsfb200 kernels
and are described by sample_factory_b200.model.ModelSpec, not by torch modules."""
from sample_factory.model.model_utils import ModelModule


class Encoder(ModelModule):
    def __init__(self, cfg):
        super().__init__(cfg)

    def get_out_size(self) -> int:
        raise NotImplementedError()

    def model_to_device(self, device):
        self.to(device)

    def device_for_input_tensor(self, input_tensor_name: str):
        return next(self.parameters()).device

    def type_for_input_tensor(self, input_tensor_name: str):
        import torch

        return torch.float32
