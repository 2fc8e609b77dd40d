"""Run the UNMODIFIED reference predator-prey model (read from /root/reference) on CPU.

TEST INFRASTRUCTURE — build container only (see run_reference.py for the import recipe).  Validates
oracle/predprey_port.py and generates the golden fixtures tests/golden/predprey_*.npz (make_golden.py).

The model is the one `models/predator_prey/main.py:1-48` runs: `Runner(config, registry)` with the substeps of
`models/predator_prey/substeps/{move,eat,hunt,grow}.py`, the `grid` network helper (`core/helpers/environment.py:6-9`)
and the bundled CSV populations (`models/predator_prey/data/*`).  The only randomness of a step is
`random.choice(possible_neighbors[idx])` (substeps/move.py:88-90), called for the agents with positive energy, predators
first (config.yaml:224-226 active_agents order).  It is replaced by an object whose `choice(seq)` consumes injected
draws: seq[(u24 * len(seq)) >> 24] with u24 = oracle.noise.move_draw(seed, step, species, agent).
"""
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from oracle import noise                                                         # noqa: E402
from oracle.ref_harness.run_reference import REFERENCE_ROOT, _prepare_import     # noqa: E402

PP_ROOT = os.path.join(REFERENCE_ROOT, "agent_torch", "models", "predator_prey")


class _InjectedRandom:
    """Replaces the `random` module inside substeps/move.py."""

    def __init__(self):
        self.queue = []

    def choice(self, seq):
        u24 = self.queue.pop(0)
        return seq[(int(u24) * len(seq)) >> 24]


def build_runner(grid="18x25", variant="config.yaml", energy_seed=0, overrides=None):
    _prepare_import()
    from agent_torch.core import Registry, Runner
    from agent_torch.core.helpers import grid_network, read_config, read_from_file
    import agent_torch.models.predator_prey.substeps as _substeps                # registers the 8 substep classes  # noqa: F401
    import agent_torch.models.predator_prey.substeps.move as move_mod
    from agent_torch.models.predator_prey.helpers.random import random_float, random_int

    config = read_config(os.path.join(PP_ROOT, variant))
    meta = config["simulation_metadata"]
    mx, my = (int(v) for v in grid.split("x"))
    n_pred = len(open(os.path.join(PP_ROOT, "data", grid, "predator-coords.csv")).read().strip().splitlines()) - 1
    n_prey = len(open(os.path.join(PP_ROOT, "data", grid, "prey-coords.csv")).read().strip().splitlines()) - 1
    meta.update(max_x=mx, max_y=my, num_predators=n_pred, num_prey=n_prey, num_grass=mx * my)
    for key, fname in (("predator_coords_file", "predator-coords.csv"), ("prey_coords_file", "prey-coords.csv"),
                       ("grass_coords_file", "grass-coords.csv"), ("grass_growth_stage_file", "growth-stage.csv"),
                       ("grass_growth_countdown_file", "growth-countdown.csv")):
        meta[key] = os.path.join(PP_ROOT, "data", grid, fname)
    config = _re_resolve(config, meta)
    # the shipped config names the street-map network (needs osmnx + network access); main.py:30 also registers
    # the grid helper, which is what the bounds/coordinates of the bundled data describe
    config["state"]["network"]["agent_agent"]["predator_prey"] = {"type": "grid", "arguments": {"shape": [mx, my]}}
    for k, v in (overrides or {}).items():
        node = config
        parts = k.split("/")
        for p in parts[:-1]:
            node = node[p]
        node[parts[-1]] = v
    registry = Registry()
    registry.register(read_from_file, "read_from_file", "initialization")
    registry.register(grid_network, "grid", key="network")
    registry.register(random_float, "random_float", "initialization")
    registry.register(random_int, "random_int", "initialization")
    torch.manual_seed(energy_seed)
    runner = Runner(config, registry)
    runner.init()
    injected = _InjectedRandom()
    move_mod.random = injected
    return runner, injected


def _re_resolve(config, meta):
    """The stand-in OmegaConf resolved ${simulation_metadata.*} at load time; re-point everything that derives
    from the sizes we changed."""
    st = config["state"]
    st["environment"]["bounds"]["value"] = [meta["max_x"], meta["max_y"]]
    # config.yaml:31 says `shape: 2`; the reference initializer needs a list there (initializer.py:89-91 calls
    # torch.zeros(size=2) otherwise), docs/config-map.yaml style `[2]` keeps the int64 value tensor as it is
    st["environment"]["bounds"]["shape"] = [2]
    for species, n in (("predator", meta["num_predators"]), ("prey", meta["num_prey"])):
        st["agents"][species]["number"] = n
        props = st["agents"][species]["properties"]
        props["coordinates"]["shape"] = [n, 2]
        props["energy"]["shape"] = [n, 1]
        props["coordinates"]["initialization_function"]["arguments"]["file_path"]["value"] = meta[f"{species}_coords_file"]
    g = st["objects"]["grass"]
    g["number"] = meta["num_grass"]
    gp = g["properties"]
    gp["coordinates"]["shape"] = [meta["num_grass"], 2]
    gp["growth_stage"]["shape"] = [meta["num_grass"], 1]
    gp["growth_countdown"]["shape"] = [meta["num_grass"], 1]
    gp["coordinates"]["initialization_function"]["arguments"]["file_path"]["value"] = meta["grass_coords_file"]
    gp["growth_stage"]["initialization_function"]["arguments"]["file_path"]["value"] = meta["grass_growth_stage_file"]
    gp["growth_countdown"]["initialization_function"]["arguments"]["file_path"]["value"] = meta["grass_growth_countdown_file"]
    return config


def snapshot(state):
    s = state
    return {
        "pred_pos": s["agents"]["predator"]["coordinates"].detach().cpu().numpy().copy(),
        "pred_energy": s["agents"]["predator"]["energy"].detach().cpu().numpy().copy(),
        "prey_pos": s["agents"]["prey"]["coordinates"].detach().cpu().numpy().copy(),
        "prey_energy": s["agents"]["prey"]["energy"].detach().cpu().numpy().copy(),
        "growth_stage": s["objects"]["grass"]["growth_stage"].detach().cpu().numpy().copy(),
        "growth_countdown": s["objects"]["grass"]["growth_countdown"].detach().cpu().numpy().copy(),
    }


def run_reference_predprey(grid="18x25", num_steps=20, seed=7, variant="config.yaml", energy_seed=0, overrides=None):
    """Returns (initial snapshot, [snapshot after step 0, 1, ...], constants, adjacency)."""
    runner, injected = build_runner(grid, variant, energy_seed, overrides)
    state = runner.state
    init = snapshot(state)
    consts = {
        "bounds": state["environment"]["bounds"].cpu().numpy().copy(),
        "pred_work": float(state["agents"]["predator"]["stride_work"].item()),
        "prey_work": float(state["agents"]["prey"]["stride_work"].item()),
        "prey_nutrition": float(state["agents"]["prey"]["nutritional_value"].item()),
        "grass_nutrition": float(state["objects"]["grass"]["nutritional_value"].item()),
        "regrowth_time": float(state["objects"]["grass"]["regrowth_time"].item()),
        "update_outputs": list(runner.config["substeps"]["0"]["transition"]["update_positions"]["output_variables"]),
    }
    adjacency = state["network"]["agent_agent"]["predator_prey"]["adjacency_matrix"].cpu().numpy().copy()
    traj = []
    for t in range(num_steps):
        st = runner.state
        q = []
        for species, name in ((0, "predator"), (1, "prey")):
            e = st["agents"][name]["energy"].reshape(-1)
            alive = torch.nonzero(e > 0).reshape(-1).numpy()
            q.extend(noise.move_draw(seed, t, species, alive).tolist())
        injected.queue = q
        runner.state["current_step"] = t
        runner.step(1)
        assert not injected.queue, "unused draws: the call order differs from the assumed one"
        traj.append(snapshot(runner.state))
    return init, traj, consts, adjacency


if __name__ == "__main__":
    init, traj, consts, adj = run_reference_predprey(num_steps=3)
    print(consts, adj.shape, adj.sum())
    for k, v in traj[-1].items():
        print(k, v.dtype, v.shape, v.reshape(-1)[:6])
