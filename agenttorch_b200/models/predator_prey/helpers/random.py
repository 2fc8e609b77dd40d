"""Initialisation helpers of the predator-prey model (reference models/predator_prey/helpers/random.py:9-41)."""
import math

import torch

from agenttorch_b200.core.registry import Registry


def _limit(v):
    return float(v.reshape(-1)[0]) if torch.is_tensor(v) else float(v)


@Registry.register_helper("random_float", "initialization")
def random_float(shape, params):
    """Uniform floats in [lower_limit, upper_limit + 1) (the reference adds 1 to include the upper limit)."""
    hi, lo = _limit(params["upper_limit"]) + 1, _limit(params["lower_limit"])
    return (hi - lo) * torch.rand(tuple(shape)) + lo


@Registry.register_helper("random_int", "initialization")
def random_int(shape, params):
    hi, lo = math.floor(_limit(params["upper_limit"]) + 1), math.floor(_limit(params["lower_limit"]))
    return torch.randint(lo, hi, tuple(shape))
