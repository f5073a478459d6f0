"""sample_factory.model.core (model/core.py:9-24): base class of custom model cores."""
from sample_factory.model.model_utils import ModelModule


class ModelCore(ModelModule):
    def __init__(self, cfg):
        super().__init__(cfg)
        self.core_output_size = -1

    def get_out_size(self) -> int:
        return self.core_output_size
