"""sample_factory.utils.attr_dict (utils/attr_dict.py): dict with attribute access."""
from sample_factory_b200.cfg import AttrDict  # noqa: F401
