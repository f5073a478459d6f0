"""sample_factory.model.decoder (model/decoder.py:9-12): base class of custom decoders."""
from sample_factory.model.model_utils import ModelModule


class Decoder(ModelModule):
    pass
