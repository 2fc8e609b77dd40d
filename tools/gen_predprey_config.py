"""Emit agenttorch_b200/models/predator_prey/yamls/config.yaml: the reference's predator-prey schema
(simulation_metadata / state{environment, agents, objects, network} / substeps '0'..'3') with a seeded synthetic
population instead of CSV files and the `grid` network."""
import os

import yaml

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
M = "${simulation_metadata.%s}"


def arg(name, value, shape=(1,), learnable=False):
    return {"name": name, "learnable": learnable, "shape": list(shape) if shape is not None else None, "value": value,
            "initialization_function": None}


def prop(name, shape, dtype, value=None, gen=None, args=None, learnable=False):
    d = {"name": name, "learnable": learnable, "shape": shape, "dtype": dtype}
    if gen is None:
        d["value"] = value
        d["initialization_function"] = None
    else:
        d["initialization_function"] = {"generator": gen, "arguments": args or {}}
    return d


def where(stream):
    return {"max_x": arg("Grid rows", M % "max_x"), "max_y": arg("Grid columns", M % "max_y"),
            "seed": arg("Population seed", M % "seed"), "stream": arg("Random stream", stream)}


def energy(lo, hi, stream):
    return {"lower_limit": arg("Lower Bound", lo), "upper_limit": arg("Upper Bound", hi),
            "seed": arg("Population seed", M % "seed"), "stream": arg("Random stream", stream)}


def animal(key, lo, hi, stream, work, extra=None):
    n = "${state.agents.%s.number}" % key
    props = {
        "coordinates": prop("Position", [n, 2], "int", gen="pp_random_coords", args=where(stream)),
        "energy": prop("Energy", [n, 1], "float", gen="pp_random_energy", args=energy(lo, hi, stream)),
        "stride_work": prop("Energy Lost in Taking One Step", [1], "float", value=work, learnable=True),
    }
    props.update(extra or {})
    return {"number": M % ("num_predators" if key == "predator" else "num_prey"), "properties": props}


G = "${state.objects.grass.number}"
grass_args = {"regrowth_time": arg("Regrowth time", 100), "grown_fraction": arg("Share of grown grass", 0.4),
              "seed": arg("Population seed", M % "seed")}


def fn(generator, inputs, outputs):
    return {"generator": generator, "arguments": None, "input_variables": inputs, "output_variables": outputs}


def neighbors(agent):
    return {"find_neighbors": fn("FindNeighbors", {"bounds": "environment/bounds",
                                                   "adj_grid": "network/agent_agent/predator_prey/adjacency_matrix",
                                                   "positions": f"agents/{agent}/coordinates"}, ["possible_neighbors"])}


def movement(agent):
    return {"decide_movement": fn("DecideMovement", {"positions": f"agents/{agent}/coordinates", "energy": f"agents/{agent}/energy"},
                                  ["next_positions"])}


cfg = {
    "simulation_metadata": {
        "device": "auto", "num_episodes": 3, "num_steps_per_episode": 20, "num_substeps_per_step": 4, "calibration": False,
        "visualize": False, "seed": 0, "max_x": 64, "max_y": 64, "num_predators": 400, "num_prey": 800, "num_grass": 4096,
        "cuda_params": {"fused_episode": True, "cuda_graph": True, "inplace_progress": False, "sort_agents_by_cell": True,
                        "resort_every": 0},
    },
    "state": {
        "environment": {"bounds": prop("Bounds", [2], "int", value=[M % "max_x", M % "max_y"])},
        "agents": {
            "predator": animal("predator", 30, 100, 0, 1),
            "prey": animal("prey", 40, 100, 1, 5, {"nutritional_value": prop("Energy Gained Upon Consuming Prey", [1], "float", value=20, learnable=True)}),
        },
        "objects": {"grass": {"number": M % "num_grass", "properties": {
            "coordinates": prop("Position", [G, 2], "int", gen="pp_grid_coords", args={"max_y": arg("Grid columns", M % "max_y")}),
            "growth_stage": prop("The Grass Growth Stage", [G, 1], "int", gen="pp_growth_stage", args=grass_args),
            "regrowth_time": prop("Time for Grass to Regrow Completely", [1], "float", value=100, learnable=True),
            "growth_countdown": prop("Time for Grass to Grow Fully", [G, 1], "float", gen="pp_growth_countdown", args=grass_args),
            "nutritional_value": prop("Energy Gained Upon Consuming Grass at Max Growth Stage", [1], "float", value=7, learnable=True),
        }}},
        "network": {"agent_agent": {"predator_prey": {"type": "grid", "arguments": {"shape": [M % "max_x", M % "max_y"]}}}},
    },
    "substeps": {
        "0": {"name": "Move", "description": "Moving", "active_agents": ["predator", "prey"],
              "observation": {"predator": neighbors("predator"), "prey": neighbors("prey")},
              "policy": {"predator": movement("predator"), "prey": movement("prey")},
              "transition": {"update_positions": fn("UpdatePositions", {
                  "prey_pos": "agents/prey/coordinates", "prey_energy": "agents/prey/energy", "pred_pos": "agents/predator/coordinates",
                  "pred_energy": "agents/predator/energy", "prey_work": "agents/prey/stride_work", "pred_work": "agents/predator/stride_work"},
                  ["prey_pos", "prey_energy", "pred_pos", "pred_energy"])}},
        "1": {"name": "Eat", "description": "Eating Grass", "active_agents": ["prey"], "observation": {"prey": None},
              "policy": {"prey": {"find_eatable_grass": fn("FindEatableGrass", {
                  "bounds": "environment/bounds", "positions": "agents/prey/coordinates", "grass_growth": "objects/grass/growth_stage"},
                  ["eatable_grass_positions"])}},
              "transition": {"eat_grass": fn("EatGrass", {
                  "energy": "agents/prey/energy", "grass_growth": "objects/grass/growth_stage", "growth_countdown": "objects/grass/growth_countdown",
                  "bounds": "environment/bounds", "prey_pos": "agents/prey/coordinates", "nutrition": "objects/grass/nutritional_value",
                  "regrowth_time": "objects/grass/regrowth_time"}, ["energy", "grass_growth", "growth_countdown"])}},
        "2": {"name": "Hunt", "description": "Hunting Prey", "active_agents": ["predator"], "observation": {"predator": None},
              "policy": {"predator": {"find_targets": fn("FindTargets", {
                  "prey_pos": "agents/prey/coordinates", "pred_pos": "agents/predator/coordinates"}, ["target_positions"])}},
              "transition": {"hunt_prey": fn("HuntPrey", {
                  "prey_energy": "agents/prey/energy", "pred_energy": "agents/predator/energy",
                  "nutritional_value": "agents/prey/nutritional_value", "prey_pos": "agents/prey/coordinates",
                  "pred_pos": "agents/predator/coordinates"}, ["prey_energy", "pred_energy"])}},
        "3": {"name": "Grow", "description": "Grow Grass", "active_agents": ["prey"], "observation": {"prey": None},
              "policy": {"prey": None},
              "transition": {"grow_grass": fn("GrowGrass", {
                  "grass_growth": "objects/grass/growth_stage", "growth_countdown": "objects/grass/growth_countdown"},
                  ["grass_growth", "growth_countdown"])}},
    },
}

out = os.path.join(ROOT, "agenttorch_b200", "models", "predator_prey", "yamls", "config.yaml")
os.makedirs(os.path.dirname(out), exist_ok=True)
with open(out, "w", encoding="utf-8") as f:
    f.write("# predator-prey grid model on the B200 path; generated by tools/gen_predprey_config.py\n")
    f.write("# (schema of the reference's models/predator_prey/config.yaml; synthetic seeded population, `grid` network)\n")
    yaml.safe_dump(cfg, f, sort_keys=False, width=120)
print("wrote", out)
