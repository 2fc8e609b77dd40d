"""predator-prey grid model: the reference's registry names (models/predator_prey/substeps/*.py, main.py:26-30)
with the B200 substep classes behind them, plus a fused multi-step path for the Runner."""
from agenttorch_b200.core.helpers import grid_network, read_from_file
from agenttorch_b200.core.registry import Registry
from agenttorch_b200.core.runner import FUSED_PLANNERS


def get_registry():
    reg = Registry()
    from . import substeps  # noqa: F401  (registers the eight substep classes)
    from . import synthetic  # noqa: F401  (registers the pp_* initialisers)
    from .helpers import random_float, random_int

    reg.register(read_from_file, "read_from_file", "initialization")
    reg.register(grid_network, "grid", key="network")
    reg.register(random_float, "random_float", "initialization")
    reg.register(random_int, "random_int", "initialization")
    return reg


registry = get_registry()

from .fused import plan_for  # noqa: E402

if plan_for not in FUSED_PLANNERS:
    FUSED_PLANNERS.append(plan_for)


def get_runner(config, registry=None):
    from agenttorch_b200.core.runner import Runner

    return Runner(config, get_registry())          # re-assert this model's helpers in the shared tables (see Executor)
