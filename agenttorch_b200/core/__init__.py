from .registry import Registry
from .runner import Runner
from .controller import Controller
from .initializer import Initializer
from .executor import Executor
from .dataloader import DataLoader, LoadPopulation
from .environment import envs
from .helpers import *  # noqa: F401,F403
from . import distributions  # noqa: F401
