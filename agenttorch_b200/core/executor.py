"""Executor (reference: agent_torch/core/executor.py:8-64): model module + population -> Runner."""
import importlib

from .dataloader import DataLoader
from .runner import Runner


class BaseExecutor:
    def __init__(self, model):
        self.model = model

    def _get_runner(self, config):
        # The registry tables are class-level and shared by every model in the process (registry.py:8-14): another
        # model imported later may have re-registered a helper of the same name (both covid and predator_prey bring a
        # `read_from_file`).  Re-assert this model's registrations before building its runner.
        module = importlib.import_module(self.model.__name__)
        registry = module.get_registry() if hasattr(module, "get_registry") else module.registry
        return Runner(config, registry)


class Executor(BaseExecutor):
    def __init__(self, model, data_loader=None, pop_loader=None) -> None:
        super().__init__(model)
        if pop_loader:
            self.pop_loader = pop_loader
            self.data_loader = DataLoader(model, pop_loader)
        else:
            self.data_loader = data_loader
        self.config = self.data_loader.get_config()
        self.runner = self._get_runner(self.config)

    def init(self):
        self.config = self.data_loader.get_config()
        self.runner = self._get_runner(self.config)
        self.runner.init()

    def execute(self, key=None):
        meta = self.config["simulation_metadata"]
        for _ in range(meta["num_episodes"]):
            self.runner.reset()
            self.runner.step(meta["num_steps_per_episode"])
        if key is not None:
            self.simulation_values = self.get_simulation_values(key)

    def get_simulation_values(self, key, key_type="environment"):
        self.simulation_values = self.runner.state_trajectory[-1][-1][key_type][key]
        return self.simulation_values
