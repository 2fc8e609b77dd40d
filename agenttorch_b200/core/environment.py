"""envs.create (reference: agent_torch/core/environment.py:73-159): model + population (+ parameters,
archetypes) -> initialised Runner.  `distributed=True` returns a DistributedRunner facade: one process per
GPU is expected to be launched by torchrun (this package never spawns processes itself)."""
from typing import Any, Dict, Optional, Type

from .dataloader import LoadPopulation
from .executor import Executor


class envs:
    @staticmethod
    def create(model: Type, population: Any, parameters: Optional[Dict[str, Any]] = None,
               archetypes: Optional[Dict[str, Any]] = None, distributed: bool = False,
               world_size: Optional[int] = None, distributed_config: Optional[Dict[str, Any]] = None) -> Any:
        loader = population if isinstance(population, LoadPopulation) else LoadPopulation(population)
        simulation = Executor(model=model, pop_loader=loader)
        runner = simulation.runner
        if distributed:
            from agenttorch_b200.distributed import DistributedRunner

            runner = DistributedRunner(runner, world_size=world_size, distributed_config=distributed_config)
        if archetypes is not None:
            table = {name: obj for cat in runner.registry.helpers.values() for name, obj in cat.items()}
            for substep_name, archetype in archetypes.items():
                cls = table.get(substep_name)
                if cls is not None and hasattr(cls, "set_behavior"):
                    # an archetype is a callable observation -> (N,1) isolation tensor (the reference wraps an LLM
                    # Behavior here, environment.py:132-146; LLM machinery is out of scope for this package)
                    cls.set_behavior(archetype)
        runner.init()
        if parameters is not None:
            runner._set_parameters(parameters)
        return runner
