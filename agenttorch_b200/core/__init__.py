from .registry import Registry
from .runner import Runner
from .controller import Controller
from .initializer import Initializer
from .executor import Executor
from .dataloader import DataLoader, LoadPopulation
from .environment import envs
from .helpers import *  # noqa: F401,F403
from .vectorization import VectorizedRunner, is_vectorized, launch_distributed_simulation, vectorized  # noqa: F401
from . import distributions  # noqa: F401
from agenttorch_b200 import __version__  # noqa: F401


def __getattr__(name):
    # DistributedRunner pulls in the CUDA engine modules: resolved on first use (reference: core/__init__.py:7)
    if name == "DistributedRunner":
        from agenttorch_b200.distributed import DistributedRunner

        return DistributedRunner
    raise AttributeError(f"module {__name__!r} has no attribute {name!r}")
