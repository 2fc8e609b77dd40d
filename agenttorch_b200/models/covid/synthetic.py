"""Config-driven synthetic populations for the covid model (BASELINE configs C2/C3/C5).

Registers initialisation / network helpers under the same Registry keys the file-based ones use, so that a
YAML (or the dict returned by `synthetic_config`) selects a seeded synthetic population exactly the way the
bundled config selects `populations/astoria`:
    age ~ astoria's age histogram; 1 % of agents start exposed; in-degree ~ Poisson(d) (or NB), sources
    uniform ("rand") or 80 % inside 64k-id communities ("comm"); edge type/weight == 1.
The network helper returns the prebuilt destination-sorted CSR as the `graph` element of its result so the
engine does not have to re-sort 1e8+ edges; the COO `(edge_index, edge_attr)` pair is still provided."""
import copy
import os

import torch
import yaml

from agenttorch_b200.core.helpers import resolve_config
from agenttorch_b200.core.registry import Registry
from agenttorch_b200.graph import (mean_interactions_from_age, synthetic_contact_graph, synthetic_population)

_POP_CACHE = {}


def _dev(params):
    return torch.device(params.get("device", "cuda" if torch.cuda.is_available() else "cpu")) if isinstance(params, dict) else None


def _population(params):
    n = int(torch.as_tensor(params["num_agents"]).reshape(-1)[0].item())
    seed = int(torch.as_tensor(params["seed"]).reshape(-1)[0].item())
    frac = float(torch.as_tensor(params.get("initial_exposed", 0.01)).reshape(-1)[0].item())
    key = (n, seed, frac)
    if key not in _POP_CACHE:
        _POP_CACHE.clear()
        dev = "cuda" if torch.cuda.is_available() else "cpu"
        _POP_CACHE[key] = synthetic_population(n, seed, initial_exposed=frac, device=dev)
    return _POP_CACHE[key]


@Registry.register_helper("synthetic_age", "initialization")
def synthetic_age(shape, params):
    return _population(params)[0].float().reshape(-1, 1)


@Registry.register_helper("synthetic_disease_stage", "initialization")
def synthetic_disease_stage(shape, params):
    return _population(params)[1].clone().reshape(-1, 1)


def _timers(stage, default, a, b):
    out = torch.full_like(stage, float(default))
    out[stage == 1] = float(a)
    out[stage == 2] = float(b)
    return out


@Registry.register_helper("synthetic_infected_time", "initialization")
def synthetic_infected_time(shape, params):
    return _timers(_population(params)[1].reshape(-1, 1), 500, -1, -3)


@Registry.register_helper("synthetic_next_stage_time", "initialization")
def synthetic_next_stage_time(shape, params):
    return _timers(_population(params)[1].reshape(-1, 1), 500, 3, 5)


@Registry.register_helper("synthetic_mean_interactions", "initialization")
def synthetic_mean_interactions(shape, params):
    return mean_interactions_from_age(_population(params)[0]).reshape(-1, 1)


@Registry.register_helper("synthetic_contact_network", "network")
def synthetic_contact_network(params):
    n, d, seed = int(params["num_agents"]), float(params["avg_degree"]), int(params["seed"])
    dev = "cuda" if torch.cuda.is_available() else "cpu"
    age, _ = synthetic_population(n, int(params.get("population_seed", seed)), device=dev)
    graph = synthetic_contact_graph(n, d, seed, family=params.get("family", "rand"), degree_dist=params.get("degree_dist", "poisson"),
                                    device=dev, mean_interactions=mean_interactions_from_age(age))
    edge_index = graph.edge_index()
    e = edge_index.shape[1]
    edge_attr = torch.ones((2, e), dtype=torch.float32, device=edge_index.device)
    return graph, (edge_index, edge_attr)


def synthetic_config(num_agents, avg_degree=20.0, seed=20261019, family="rand", degree_dist="poisson",
                     initial_exposed=0.01, num_steps=21, device="auto", noise_seed=None):
    """Resolved config dict: the packaged covid YAML with the population/network sections pointed at the
    synthetic helpers above."""
    here = os.path.dirname(os.path.abspath(__file__))
    with open(os.path.join(here, "yamls", "config.yaml"), "r", encoding="utf-8") as f:
        cfg = yaml.safe_load(f)
    meta = cfg["simulation_metadata"]
    meta["num_agents"] = int(num_agents)
    meta["num_steps_per_episode"] = int(num_steps)
    meta["device"] = device
    if noise_seed is not None:
        meta["noise_seed"] = int(noise_seed)

    def a(value, shape=(1,)):
        return {"name": "", "shape": list(shape), "learnable": False, "initialization_function": None, "value": value}

    pop_args = {"num_agents": a(int(num_agents)), "seed": a(int(seed)), "initial_exposed": a(float(initial_exposed))}
    props = cfg["state"]["agents"]["citizens"]["properties"]
    for prop, gen in (("age", "synthetic_age"), ("disease_stage", "synthetic_disease_stage"),
                      ("infected_time", "synthetic_infected_time"), ("next_stage_time", "synthetic_next_stage_time")):
        props[prop]["initialization_function"] = {"generator": gen, "arguments": copy.deepcopy(pop_args)}
    del props["id"]
    cfg["state"]["environment"]["mean_interactions"]["initialization_function"] = {
        "generator": "synthetic_mean_interactions", "arguments": copy.deepcopy(pop_args)}
    cfg["state"]["network"]["agent_agent"]["infection_network"] = {
        "type": "synthetic_contact_network",
        "arguments": {"num_agents": int(num_agents), "avg_degree": float(avg_degree), "seed": int(seed) + 1,
                      "population_seed": int(seed), "family": family, "degree_dist": degree_dist}}
    return resolve_config(cfg)
