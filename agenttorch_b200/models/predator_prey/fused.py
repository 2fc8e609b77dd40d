"""Fused execution of the predator-prey config: `runner.step(n)` = ONE captured CUDA graph of 1 + 2n kernels
instead of n x 4 substeps x (observe, act, progress) Python passes (reference core/runner.py:121-153).

The plan is recognised from the initialised substep modules (exactly this package's classes under the reference's
registry names); anything else runs through the generic Controller path.  Positions, energies and growth stages
are converted at the boundary: the engine keeps the agents as (cell, energy) pairs sorted by cell (locality of the
per-cell lookups), the public tensors stay in agent-number order and are updated in place after every call."""
import re

import torch

from agenttorch_b200.core.helpers import get_by_path, set_by_path
from agenttorch_b200.predprey import PredPreyEngine, PredPreySpec, PredPreyState, unpack_cells, unpack_stage

from .context import topology


def _path(p):
    return re.split("/", p)


def plan_for(runner):
    """Returns a callable `step(num_steps)` when the runner's config is the predator-prey pattern, else None."""
    cached = getattr(runner, "_pp_plan", None)
    if cached is not None:
        return cached or None
    runner._pp_plan = False
    from .substeps.eat import EatGrass, FindEatableGrass
    from .substeps.grow import GrowGrass
    from .substeps.hunt import FindTargets, HuntPrey
    from .substeps.move import DecideMovement, FindNeighbors, UpdatePositions

    init = runner.initializer
    if init.device.type != "cuda":
        return None
    subs = runner.config["substeps"]
    tf, pf, of = init.transition_function, init.policy_function, init.observation_function
    want_t = [("update_positions", UpdatePositions), ("eat_grass", EatGrass), ("hunt_prey", HuntPrey), ("grow_grass", GrowGrass)]
    if list(subs.keys()) != ["0", "1", "2", "3"]:
        return None
    for s, (name, cls) in zip(subs.keys(), want_t):
        if list(tf[s].keys()) != [name] or type(tf[s][name]) is not cls:
            return None

    def only(table, s, agent, name, cls):
        mods = table[s][agent] if agent in table[s] else {}
        return list(mods.keys()) == [name] and type(mods[name]) is cls

    if sorted(subs["0"]["active_agents"]) != ["predator", "prey"]:
        return None
    for agent in ("predator", "prey"):
        if not only(of, "0", agent, "find_neighbors", FindNeighbors) or not only(pf, "0", agent, "decide_movement", DecideMovement):
            return None
    if not only(pf, "1", "prey", "find_eatable_grass", FindEatableGrass) or not only(pf, "2", "predator", "find_targets", FindTargets):
        return None
    for s in ("1", "2", "3"):
        if any(len(of[s][a]) for a in subs[s]["active_agents"]):
            return None
    if any(len(pf["3"][a]) for a in subs["3"]["active_agents"]):
        return None
    runner._pp_plan = _FusedPredPrey(runner)
    return runner._pp_plan


class _FusedPredPrey:
    def __init__(self, runner):
        self.runner = runner
        init = runner.initializer
        self.up = init.transition_function["0"]["update_positions"]
        self.eat = init.transition_function["1"]["eat_grass"]
        self.hunt = init.transition_function["2"]["hunt_prey"]
        self.grow = init.transition_function["3"]["grow_grass"]
        self.nb = init.observation_function["0"]["prey"]["find_neighbors"]
        self.movers = [init.policy_function["0"][a]["decide_movement"] for a in ("predator", "prey")]
        self.seed = self.movers[0].seed
        self.engine = None
        self.st = None
        self._seen = None
        cuda_params = runner.config["simulation_metadata"].get("cuda_params", {}) or {}
        self.sort = bool(cuda_params.get("sort_agents_by_cell", True))
        # measured (tools/pp_probe.py): re-sorting costs more than the locality it restores, even over 100 steps
        self.resort_every = int(cuda_params.get("resort_every", 0)) if self.sort else 0

    def _tensors(self, state):
        u, e = self.up.input_variables, self.eat.input_variables
        g = lambda p: get_by_path(state, _path(p))                                      # noqa: E731
        return {"pred_pos": g(u["pred_pos"]), "pred_energy": g(u["pred_energy"]), "prey_pos": g(u["prey_pos"]),
                "prey_energy": g(u["prey_energy"]), "stage": g(e["grass_growth"]), "countdown": g(e["growth_countdown"])}

    def _build(self, state, t):
        u, e, h = self.up.input_variables, self.eat.input_variables, self.hunt.input_variables
        g = lambda p: get_by_path(state, _path(p))                                      # noqa: E731
        one = lambda p: float(g(p).reshape(-1)[0])                                      # noqa: E731
        bounds = g(e["bounds"]).reshape(-1).tolist()
        adj = g(self.nb.input_variables["adj_grid"])
        dev = t["pred_energy"].device
        lattice, csr = topology(adj, dev)
        n_cells = t["stage"].reshape(-1).shape[0]
        apply_pred, apply_prey = PredPreySpec.work_flags(self.up.output_variables)
        spec = PredPreySpec(max_x=int(bounds[0]), max_y=int(bounds[1]), n_cells=n_cells,
                            width=int(lattice.width) if lattice is not None else int(bounds[0]),
                            pred_work=one(u["pred_work"]), prey_work=one(u["prey_work"]), prey_nutrition=one(h["nutritional_value"]),
                            grass_nutrition=one(e["nutrition"]), regrowth_time=one(e["regrowth_time"]),
                            apply_pred_work=apply_pred, apply_prey_work=apply_prey,
                            tiled=self.sort and lattice is not None and n_cells % int(lattice.width) == 0)
        if lattice is not None and lattice.n_cells != n_cells:
            raise ValueError(f"grid network has {lattice.n_cells} cells but the grass has {n_cells}")
        self.engine = PredPreyEngine(spec, dev, adjacency=csr)
        self.spec = spec

    def _unchanged(self, *tensors):
        """True when the public coordinate / stage tensors are the ones this path wrote last and no torch op has
        modified them since (then the engine's packed copies are still current)."""
        return self._seen is not None and all(x is y and x._version == v for x, (y, v) in zip(tensors, self._seen))

    def _adopt(self, state):
        """Bring the engine state in line with the public tensors.  The engine keeps its own copy: agents sorted by cell,
        grass in tile order."""
        t = self._tensors(state)
        if self.engine is None:
            self._build(state, t)
        spec = self.spec
        paths_u, paths_e = self.up.input_variables, self.eat.input_variables
        fixed = {}
        for key, path, dtype in (("pred_energy", paths_u["pred_energy"], torch.float32), ("prey_energy", paths_u["prey_energy"], torch.float32),
                                 ("pred_pos", paths_u["pred_pos"], torch.int64), ("prey_pos", paths_u["prey_pos"], torch.int64)):
            v = t[key]
            if v.dtype != dtype or not v.is_contiguous():
                v = v.to(dtype).contiguous()
                set_by_path(state, _path(path), v)
            fixed[key] = v
        cd = t["countdown"]
        if cd.dtype not in (torch.float32, torch.float64) or not cd.is_contiguous():
            cd = cd.to(torch.float32).contiguous()
            set_by_path(state, _path(paths_e["growth_countdown"]), cd)
        stage = t["stage"]
        if stage.dtype not in (torch.int64, torch.float32) or not stage.is_contiguous():
            stage = stage.to(torch.int64).contiguous()
            set_by_path(state, _path(paths_e["grass_growth"]), stage)
        fixed["stage"], fixed["countdown"] = stage, cd
        watched = [fixed[k] for k in ("pred_pos", "pred_energy", "prey_pos", "prey_energy", "stage", "countdown")]
        if self.st is None or not self._unchanged(*watched):
            fresh = PredPreyState.from_reference(spec, fixed["pred_pos"], fixed["pred_energy"], fixed["prey_pos"], fixed["prey_energy"],
                                                 stage, cd, sort=self.sort)
            if self.st is not None and self.st.pred_cell.shape == fresh.pred_cell.shape and self.st.prey_cell.shape == fresh.prey_cell.shape \
                    and (self.st.pred_id is None) == (fresh.pred_id is None) and self.st.countdown.dtype == fresh.countdown.dtype:
                self.st.copy_(fresh)                                                    # keep the buffers captured graphs point to
            else:
                self.st = fresh
        return fixed

    def __call__(self, num_steps):
        runner, state = self.runner, self.runner.state
        fixed = self._adopt(state)
        st = self.st
        t0 = self.movers[0].calls
        before = self.engine.launch_count
        self.engine.run(st, int(num_steps), self.seed, t0=t0, cuda_graph=runner._use_cuda_graph, resort_every=self.resort_every)
        for m in self.movers:
            m.calls = t0 + int(num_steps)
        # publish: agents back in agent-number order, in place in the public tensors
        unpack_cells(st.in_agent_order("pred_cell"), self.spec.max_y, out=fixed["pred_pos"])
        unpack_cells(st.in_agent_order("prey_cell"), self.spec.max_y, out=fixed["prey_pos"])
        fixed["pred_energy"].view(-1).copy_(st.in_agent_order("pred_energy"))
        fixed["prey_energy"].view(-1).copy_(st.in_agent_order("prey_energy"))
        info, countdown = st.grass_in_cell_order(self.spec)
        unpack_stage(info, self.spec.n_cells, out=fixed["stage"].view(-1, 1))
        fixed["countdown"].view(-1).copy_(countdown)
        self._seen = [(x, x._version) for x in (fixed[k] for k in ("pred_pos", "pred_energy", "prey_pos", "prey_energy", "stage", "countdown"))]
        state["current_step"] = int(num_steps) - 1
        state["current_substep"] = "0"
        runner.perf_stats["fused_episodes"] += 1
        runner.perf_stats["kernel_launches"] += self.engine.launch_count - before + 3
        n_sub = len(runner.config["substeps"])
        for t in range(int(num_steps)):
            runner.state_trajectory.append([{"current_step": t, "current_substep": str((s + 1) % n_sub),
                                             "environment": dict(state["environment"])} for s in range(n_sub)])
