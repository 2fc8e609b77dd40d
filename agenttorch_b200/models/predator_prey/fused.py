"""Fused execution of the predator-prey config: `runner.step(n)` = ONE captured CUDA graph of 1 + 2n kernels
instead of n x 4 substeps x (observe, act, progress) Python passes (reference core/runner.py:121-153).

The plan is recognised from the initialised substep modules (exactly this package's classes under the reference's
registry names); anything else runs through the generic Controller path.  Energies and countdowns of the public state
are aliased by the engine (no copies); coordinates and growth stages are converted at the boundary, in place."""
import re

import torch

from agenttorch_b200.core.helpers import get_by_path, set_by_path
from agenttorch_b200.predprey import (PredPreyEngine, PredPreySpec, PredPreyState, pack_cells, pack_stage,
                                      unpack_cells, unpack_stage)

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
                            apply_pred_work=apply_pred, apply_prey_work=apply_prey)
        if lattice is not None and lattice.n_cells != n_cells:
            raise ValueError(f"grid network has {lattice.n_cells} cells but the grass has {n_cells}")
        self.engine = PredPreyEngine(spec, dev, adjacency=csr)
        self.spec = spec

    def _unchanged(self, *tensors):
        """True when the public coordinate / stage tensors are the ones this path wrote last and no torch op has
        modified them since (then the engine's packed copies are still current)."""
        return self._seen is not None and all(x is y and x._version == v for x, (y, v) in zip(tensors, self._seen))

    def _adopt(self, state):
        """Bring the engine state in line with the public tensors; aliases what can be aliased."""
        t = self._tensors(state)
        if self.engine is None:
            self._build(state, t)
        spec = self.spec
        paths_u, paths_e = self.up.input_variables, self.eat.input_variables
        fixed = {}
        for key, path in (("pred_energy", paths_u["pred_energy"]), ("prey_energy", paths_u["prey_energy"])):
            v = t[key]
            if v.dtype != torch.float32 or not v.is_contiguous():
                v = v.to(torch.float32).contiguous()
                set_by_path(state, _path(path), v)
            fixed[key] = v
        cd = t["countdown"]
        if cd.dtype not in (torch.float32, torch.float64) or not cd.is_contiguous():
            cd = cd.to(torch.float32).contiguous()
            set_by_path(state, _path(paths_e["growth_countdown"]), cd)
        for key, path in (("pred_pos", paths_u["pred_pos"]), ("prey_pos", paths_u["prey_pos"])):
            v = t[key]
            if v.dtype != torch.int64 or not v.is_contiguous():
                v = v.to(torch.int64).contiguous()
                set_by_path(state, _path(path), v)
            fixed[key] = v
        stage = t["stage"]
        if stage.dtype not in (torch.int64, torch.float32) or not stage.is_contiguous():
            stage = stage.to(torch.int64).contiguous()
            set_by_path(state, _path(paths_e["grass_growth"]), stage)
        if self.st is None or not self._unchanged(fixed["pred_pos"], fixed["prey_pos"], stage):
            pc, f0 = pack_cells(fixed["pred_pos"], spec.max_y, spec.n_cells)
            yc, f1 = pack_cells(fixed["prey_pos"], spec.max_y, spec.n_cells)
            if int((f0 | f1).item()):
                raise ValueError("agent coordinates outside the grid (cell = max_y * x + y must be in [0, n_cells))")
            info = pack_stage(stage, spec.n_cells)
            if self.st is not None and self.st.pred_cell.shape == pc.shape and self.st.prey_cell.shape == yc.shape:
                self.st.pred_cell.copy_(pc); self.st.prey_cell.copy_(yc); self.st.info.copy_(info)       # keep graph pointers
            else:
                self.st = PredPreyState(pc, None, yc, None, info, None)
        self.st.pred_energy, self.st.prey_energy = fixed["pred_energy"].view(-1), fixed["prey_energy"].view(-1)
        self.st.countdown = cd.view(-1)
        return fixed, stage

    def __call__(self, num_steps):
        runner, state = self.runner, self.runner.state
        fixed, stage = self._adopt(state)
        t0 = self.movers[0].calls
        before = self.engine.launch_count
        self.engine.run(self.st, int(num_steps), self.seed, t0=t0, cuda_graph=runner._use_cuda_graph)
        for m in self.movers:
            m.calls = t0 + int(num_steps)
        unpack_cells(self.st.pred_cell, self.spec.max_y, out=fixed["pred_pos"])
        unpack_cells(self.st.prey_cell, self.spec.max_y, out=fixed["prey_pos"])
        unpack_stage(self.st.info, self.spec.n_cells, out=stage.view(-1, 1))
        self._seen = [(x, x._version) for x in (fixed["pred_pos"], fixed["prey_pos"], stage)]
        state["current_step"] = int(num_steps) - 1
        state["current_substep"] = "0"
        runner.perf_stats["fused_episodes"] += 1
        runner.perf_stats["kernel_launches"] += self.engine.launch_count - before + 3
        n_sub = len(runner.config["substeps"])
        for t in range(int(num_steps)):
            runner.state_trajectory.append([{"current_step": t, "current_substep": str((s + 1) % n_sub),
                                             "environment": dict(state["environment"])} for s in range(n_sub)])
