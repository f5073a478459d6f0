"""sample_factory.utils.typing (utils/typing.py): the aliases user code annotates with."""
from __future__ import annotations

import argparse
from typing import Any, Callable, Dict, Optional, Tuple, Union

import torch

from sample_factory.utils.attr_dict import AttrDict

Config = Union[argparse.Namespace, AttrDict]
StatusCode = int
PolicyID = int
Device = str
MpQueue = Any
MpLock = Any
Env = Any
ObsSpace = Any
ActionSpace = Any
CreateEnvFunc = Callable[[str, Optional[Config], Optional[AttrDict], Optional[str]], Env]
ActionDistribution = Any
InitModelData = Tuple[PolicyID, Dict, torch.device, int]
