"""sample_factory.launcher.run_description (run_description.py): the DATA side of the experiment launcher -- parameter grids and
the command lines they expand to -- so that experiment definition files (sf_examples/*/experiments/*.py) import and can be
expanded; the process / slurm / ngc back ends that execute the commands are out of scope."""
from __future__ import annotations

import itertools
from os.path import join
from typing import Dict, Iterator, List, Sequence, Tuple


class ParamGenerator:
    def generate_params(self, randomize: bool = True) -> Iterator[Dict]:
        raise NotImplementedError


class ParamList(ParamGenerator):
    """explicit list of parameter combinations"""

    def __init__(self, combinations: Sequence[Dict]):
        self.combinations = list(combinations)

    def generate_params(self, randomize: bool = True) -> Iterator[Dict]:
        yield from self.combinations


class ParamGrid(ParamGenerator):
    """cartesian product of [(name, [values...]), ...]; a tuple of names walks tuples of values together"""

    def __init__(self, grid_tuples: Sequence[Tuple]):
        self.grid = list(grid_tuples)

    def generate_params(self, randomize: bool = False) -> Iterator[Dict]:
        if not self.grid:
            yield dict()
            return
        names = [g[0] for g in self.grid]
        for combo in itertools.product(*[g[1] for g in self.grid]):
            params = dict()
            for name, value in zip(names, combo):
                if isinstance(name, (tuple, list)):
                    params.update(dict(zip(name, value)))
                else:
                    params[name] = value
            yield params


class Experiment:
    def __init__(self, name: str, cmd: str, param_generator=None, env_vars: Dict | None = None):
        self.base_name, self.cmd, self.env_vars = name, cmd, env_vars
        self.params = list(param_generator) if param_generator is not None else [dict()]

    def generate_experiments(self, experiment_arg_name: str = "--experiment", customize_experiment_name: bool = True,
                             param_prefix: str = "--") -> Iterator[Tuple[str, str]]:
        """(command line, experiment name) per parameter combination"""
        for idx, combo in enumerate(self.params):
            tokens, name = [self.cmd], f"{idx:02d}_{self.base_name}"
            for key, value in combo.items():
                tokens.append(f"{param_prefix}{key}={value}")
                if customize_experiment_name:
                    short = "".join(t[0] for t in str(key).split("_"))
                    name += f"_{short}_{value}"
            tokens.append(f"{experiment_arg_name}={name}")
            yield " ".join(tokens), name


class RunDescription:
    def __init__(self, run_name: str, experiments: List[Experiment], experiment_arg_name: str = "--experiment",
                 experiment_dir_arg_name: str = "--train_dir", customize_experiment_name: bool = True, param_prefix: str = "--"):
        self.run_name, self.experiments = run_name, experiments
        self.experiment_arg_name, self.experiment_dir_arg_name = experiment_arg_name, experiment_dir_arg_name
        self.customize_experiment_name, self.param_prefix = customize_experiment_name, param_prefix

    def generate_experiments(self, train_dir: str, makedirs: bool = False) -> Iterator[Tuple[str, str, str, Dict | None]]:
        """(command line, experiment name, root dir, env vars) for every experiment of the run"""
        for experiment in self.experiments:
            root_dir = join(self.run_name, experiment.base_name)
            for cmd, name in experiment.generate_experiments(self.experiment_arg_name, self.customize_experiment_name,
                                                             self.param_prefix):
                yield f"{cmd} {self.experiment_dir_arg_name}={join(train_dir, root_dir)}", name, root_dir, experiment.env_vars
