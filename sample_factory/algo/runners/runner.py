"""sample_factory.algo.runners.runner (runner.py:52-184): the observer interface and the runner class of the device engine."""
from sample_factory_b200.train import Runner  # noqa: F401


class AlgoObserver:
    """hooks the runner calls (runner.py:52-73); every one receives the runner first"""

    def on_init(self, runner) -> None:
        pass

    def on_connect_components(self, runner) -> None:
        """(the device engine has no signal-slot graph to extend: never called)"""

    def on_start(self, runner) -> None:
        pass

    def on_training_step(self, runner, training_iteration_since_resume: int) -> None:
        pass

    def extra_summaries(self, runner, policy_id, env_steps: int, writer) -> None:
        pass

    def on_stop(self, runner) -> None:
        pass
