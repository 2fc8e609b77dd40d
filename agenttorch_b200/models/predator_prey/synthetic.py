"""Seeded synthetic populations for the predator-prey model (BASELINE configs[3]: 10 M agents on a grid).

Registered as initialisation helpers, so a YAML selects them the way the reference's config selects its CSV files
(`read_from_file`, config.yaml:47-55).  Everything is generated on the device from a seed: uniform positions,
uniform energies, and a grass field whose stage is consistent with its countdown (stage 1 <=> countdown <= 0)."""
import copy
import os

import torch
import yaml

from agenttorch_b200.core.helpers import resolve_config
from agenttorch_b200.core.registry import Registry


def _val(params, key, default=None):
    v = params.get(key, default)
    return float(torch.as_tensor(v).reshape(-1)[0].item())


def _gen(params, salt):
    dev = "cuda" if torch.cuda.is_available() else "cpu"
    g = torch.Generator(device=dev)
    g.manual_seed(int(_val(params, "seed", 0)) * 1000003 + salt)
    return g, dev


@Registry.register_helper("pp_random_coords", "initialization")
def pp_random_coords(shape, params):
    g, dev = _gen(params, int(_val(params, "stream", 0)))
    n = int(shape[0])
    x = torch.randint(0, int(_val(params, "max_x")), (n,), generator=g, device=dev)
    y = torch.randint(0, int(_val(params, "max_y")), (n,), generator=g, device=dev)
    return torch.stack([x, y], dim=1)


@Registry.register_helper("pp_random_energy", "initialization")
def pp_random_energy(shape, params):
    g, dev = _gen(params, 100 + int(_val(params, "stream", 0)))
    lo, hi = _val(params, "lower_limit"), _val(params, "upper_limit") + 1          # helpers/random.py:18-19
    return (hi - lo) * torch.rand(tuple(shape), generator=g, device=dev) + lo


def _grass(shape, params):
    g, dev = _gen(params, 200)
    n = int(shape[0])
    grown = torch.rand(n, generator=g, device=dev) < _val(params, "grown_fraction", 0.4)
    left = torch.randint(1, max(2, int(_val(params, "regrowth_time", 100))) + 1, (n,), generator=g, device=dev)
    return torch.where(grown, torch.zeros_like(left), left)


@Registry.register_helper("pp_growth_countdown", "initialization")
def pp_growth_countdown(shape, params):
    return _grass(shape, params).to(torch.float32).reshape(-1, 1)


@Registry.register_helper("pp_growth_stage", "initialization")
def pp_growth_stage(shape, params):
    return (_grass(shape, params) <= 0).to(torch.int64).reshape(-1, 1)


@Registry.register_helper("pp_grid_coords", "initialization")
def pp_grid_coords(shape, params):
    dev = "cuda" if torch.cuda.is_available() else "cpu"
    my = int(_val(params, "max_y"))
    k = torch.arange(int(shape[0]), device=dev)
    return torch.stack([k // my, k % my], dim=1)


def default_config_path():
    return os.path.join(os.path.dirname(os.path.abspath(__file__)), "yamls", "config.yaml")


def synthetic_config(max_x, max_y, num_predators, num_prey, seed=0, num_steps=20, device="cuda", grass_coordinates=True, **meta):
    """The packaged config re-sized (the same edit a user would make in the YAML)."""
    with open(default_config_path(), "r", encoding="utf-8") as f:
        raw = yaml.safe_load(f)
    raw = copy.deepcopy(raw)
    raw["simulation_metadata"].update(max_x=int(max_x), max_y=int(max_y), num_predators=int(num_predators), num_prey=int(num_prey),
                                      num_grass=int(max_x) * int(max_y), seed=int(seed), num_steps_per_episode=int(num_steps),
                                      device=device, **meta)
    if not grass_coordinates:                      # (cells,2) int64 that no substep reads: 268 MB at 4096 x 4096
        del raw["state"]["objects"]["grass"]["properties"]["coordinates"]
    return resolve_config(raw)
