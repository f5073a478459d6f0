"""sample_factory.cfg.arguments (cfg/arguments.py:24-224): the two-pass parser and the cfg helpers."""
from sample_factory_b200.cfg import (  # noqa: F401
    default_cfg,
    parse_full_cfg,
    parse_sf_args,
    preprocess_cfg,
    verify_cfg,
)
from sample_factory_b200.enjoy import checkpoint_override_defaults, load_from_checkpoint  # noqa: F401
