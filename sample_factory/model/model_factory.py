"""sample_factory.model.model_factory (model/model_factory.py:16-60)."""
from sample_factory_b200.model_factory import ModelFactory, UnsupportedCustomModel  # noqa: F401
