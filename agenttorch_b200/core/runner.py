"""Runner — episode loop with the reference's public API (agent_torch/core/runner.py:10-594):
init / reset / reset_state / step(n) / step_from_params / forward / _set_parameters /
get_performance_stats / state / state_trajectory.

Two execution paths with identical results:
  * generic: observe -> act -> progress per substep through the Controller (any registered substeps);
  * fused episode: when the config is the SEIRM pattern served by this package's substep classes
    (policy MakeIsolationDecision without a behavior, transitions NewTransmission + SEIRMProgression),
    `step(n)` issues ONE differentiable operator (`agenttorch_b200::seirm_episode`) that advances the state
    tensors in place with one fused CUDA launch per step, replayed from a CUDA graph — no per-substep
    state copies (controller.py:64), no host round trips.
Trajectory snapshots hold the environment only (as the reference's CUDA path, runner.py:224-266) and,
unlike it, stay attached to the autograd graph."""
import torch
import torch.nn as nn

from .controller import Controller
from .initializer import Initializer


# callables `(runner) -> callable(num_steps) | None`, appended by the model packages that ship a fused path
FUSED_PLANNERS = []


class _LazyEnvironment(dict):
    """Environment view of an intermediate (step, substep) of a fused episode: the (T,) series are masked to
    the steps already simulated when they are read."""

    def __init__(self, env, upto):
        super().__init__(env)
        self._upto = upto

    def __getitem__(self, key):
        v = super().__getitem__(key)
        if key in self._upto and torch.is_tensor(v):
            n = self._upto[key]
            mask = (torch.arange(v.shape[-1], device=v.device) < n).to(v.dtype)
            return v * mask
        return v


class Runner(nn.Module):
    def __init__(self, config, registry) -> None:
        super().__init__()
        self.config = config
        self.registry = registry
        meta = self.config["simulation_metadata"]
        assert meta["num_substeps_per_step"] == len(list(self.config["substeps"].keys()))
        cuda_params = meta.get("cuda_params", {}) or {}
        self._fused_enabled = bool(cuda_params.get("fused_episode", True))
        self._use_cuda_graph = bool(cuda_params.get("cuda_graph", True))
        self._inplace_progress = bool(cuda_params.get("inplace_progress", False))
        self.use_gpu = torch.cuda.is_available()
        self.device = torch.device("cuda" if self.use_gpu else "cpu")
        self.perf_stats = {"gpu_to_cpu_transfers": 0, "tensor_allocations": 0, "memory_reused": 0,
                           "vectorized_operations": 0, "fused_episodes": 0, "kernel_launches": 0}
        self.initializer = Initializer(self.config, self.registry)
        self.controller = Controller(self.config)
        self.state = None
        self.state_trajectory = []
        self._episode_index = 0

    # ------------------------------------------------------------------------------ lifecycle
    def _snapshot(self, state):
        """Environment-only snapshot; tensors are not copied (they are replaced, never mutated, by progress)."""
        return {"current_step": int(state.get("current_step", 0)), "current_substep": state.get("current_substep", "0"),
                "environment": dict(state.get("environment", {}))}

    def init(self):
        self._plan_cache = None               # the substep modules are rebuilt: fused plans must be re-derived
        self._pp_plan = None
        self.initializer.initialize()
        self.state = self.initializer.state
        self.state_trajectory = [[self._snapshot(self.state)]]

    def reset(self):
        self.init()

    def reset_state(self):
        self.initializer.reset_state()
        self.state = self.initializer.state
        self.state_trajectory = [[self._snapshot(self.state)]]

    def restore_state(self, snapshot, non_blocking=True):
        """Copy a saved agent/environment state (nested dict of tensors, host or device; e.g. pinned host
        buffers of an initial condition or a checkpoint, possibly in narrower dtypes) INTO the live state tensors, keeping their storage —
        and therefore the engine bindings and captured CUDA graphs — valid.  The cheap counterpart of
        reset_state(), which re-runs every initialiser."""
        from .snapshot import PackedSnapshot

        if isinstance(snapshot, PackedSnapshot):
            # one pinned host buffer: one copy over PCIe, widened on the device
            snapshot.restore_into(self.state, non_blocking=non_blocking)
            self.state["current_step"] = 0
            self.state["current_substep"] = "0"
            self.state_trajectory = [[self._snapshot(self.state)]]
            return

        def walk(dst, src):
            for k, v in src.items():
                if isinstance(v, dict):
                    walk(dst[k], v)
                elif torch.is_tensor(v) and torch.is_tensor(dst.get(k)):
                    cur = dst[k]
                    if cur.requires_grad or cur.grad_fn is not None or cur.shape != v.shape:
                        dst[k] = v.to(cur.device, non_blocking=non_blocking).reshape(cur.shape).to(cur.dtype)
                    elif v.dtype != cur.dtype and v.device != cur.device:
                        # compact snapshots (e.g. uint8 stages, int16 timers): move the narrow bytes, widen on the device
                        cur.copy_(v.to(cur.device, non_blocking=non_blocking))
                    else:
                        cur.copy_(v, non_blocking=non_blocking)
        walk(self.state, snapshot)
        self.state["current_step"] = 0
        self.state["current_substep"] = "0"
        self.state_trajectory = [[self._snapshot(self.state)]]

    # ---------------------------------------------------------------------------- checkpoint / resume
    def _noise_modules(self):
        """(name, module) of every substep module that owns a counter-hash noise address."""
        init = self.initializer
        for kind, table in (("observation", init.observation_function), ("policy", init.policy_function)):
            for substep, per_agent in table.items():
                for agent, mods in per_agent.items():
                    for name, mod in mods.items():
                        if hasattr(mod, "noise_seed"):
                            yield f"{kind}.{substep}.{agent}.{name}", mod
        for substep, mods in init.transition_function.items():
            for name, mod in mods.items():
                if hasattr(mod, "noise_seed"):
                    yield f"transition.{substep}.{name}", mod

    def checkpoint(self):
        """Everything a later episode depends on, as CPU tensors and plain numbers: the agent / environment / object state
        (SoA tensors), the learnable and calibration parameters of every substep module, and the noise counters (seed, episode
        index) of the modules that draw from the counter hash — the reference pickles `state` + `parameters_dict`
        (initializer.py:434-444) and has no reproducible random stream to save.  Episodes run from a restored checkpoint
        are bitwise identical to the ones the original runner would have run."""
        def to_cpu_tree(tree):
            out = {}
            for k, v in tree.items():
                if isinstance(v, dict):
                    out[k] = to_cpu_tree(v)
                elif torch.is_tensor(v):
                    out[k] = v.detach().to("cpu").clone()
            return out

        params = {}
        for mod_name, mod in self.initializer.named_modules():
            for attr, value in vars(mod).items():
                if attr.startswith("calibrate_") and torch.is_tensor(value):
                    params[f"{mod_name}.{attr}"] = value.detach().to("cpu").clone()
            if hasattr(mod, "external_M") and torch.is_tensor(mod.external_M):
                params[f"{mod_name}.external_M"] = mod.external_M.detach().to("cpu").clone()
        noise = {name: {"noise_seed": int(m.noise_seed), "episode_index": int(m.episode_index), "advance_seed": bool(m.advance_seed)}
                 for name, m in self._noise_modules()}
        return {"format": 1, "num_agents": int(self.config["simulation_metadata"]["num_agents"]),
                "state": to_cpu_tree({k: v for k, v in self.state.items() if k in ("agents", "environment", "objects")}),
                "module_state": {k: v.detach().to("cpu").clone() for k, v in self.initializer.state_dict().items()},
                "parameters": params, "noise": noise, "episode_index": int(self._episode_index)}

    def save_checkpoint(self, path):
        torch.save(self.checkpoint(), path)

    def load_checkpoint(self, source):
        """Restore a checkpoint (a path or the dict `checkpoint()` returned) into this initialised runner."""
        ck = torch.load(source, map_location="cpu", weights_only=True) if isinstance(source, (str, bytes)) or hasattr(source, "__fspath__") else source
        if ck.get("format") != 1:
            raise ValueError(f"unsupported checkpoint format {ck.get('format')}")
        if ck["num_agents"] != int(self.config["simulation_metadata"]["num_agents"]):
            raise ValueError(f"checkpoint of {ck['num_agents']} agents, runner has {self.config['simulation_metadata']['num_agents']}")
        self.restore_state(ck["state"], non_blocking=False)
        self.initializer.load_state_dict(ck["module_state"], strict=False)
        modules = dict(self.initializer.named_modules())
        for key, value in ck["parameters"].items():
            mod_name, attr = key.rsplit(".", 1)
            mod = modules[mod_name]
            cur = getattr(mod, attr)
            setattr(mod, attr, value.to(cur.device).to(cur.dtype).requires_grad_(cur.requires_grad))
        mine = dict(self._noise_modules())
        for name, counters in ck["noise"].items():
            if name in mine:
                mod = mine[name]
                mod.noise_seed, mod.episode_index, mod.advance_seed = counters["noise_seed"], counters["episode_index"], counters["advance_seed"]
        self._episode_index = int(ck.get("episode_index", 0))

    def reset_state_before_episode(self):
        self.state_trajectory = [[self._snapshot(self.state)]]

    # ---------------------------------------------------------------------------------- step
    def step(self, num_steps=None):
        assert self.state is not None
        if not num_steps:
            num_steps = self.config["simulation_metadata"]["num_steps_per_episode"]
        plan = self._fused_plan() if self._fused_enabled else None
        if plan is not None:
            return self._step_fused(num_steps, plan)
        if self._fused_enabled:
            for planner in FUSED_PLANNERS:     # other models served by one fused launch sequence per step() call
                fused_step = planner(self)
                if fused_step is not None:
                    return fused_step(num_steps)
        return self._step_generic(num_steps)

    def _step_generic(self, num_steps):
        for time_step in range(num_steps):
            self.state["current_step"] = time_step
            self.state_trajectory.append([])
            for substep in self.config["substeps"].keys():
                observation_profile, action_profile = {}, {}
                for agent_type in self.config["substeps"][substep]["active_agents"]:
                    assert substep == self.state["current_substep"]
                    observation_profile[agent_type] = self.controller.observe(
                        self.state, self.initializer.observation_function, agent_type)
                    action_profile[agent_type] = self.controller.act(
                        self.state, observation_profile[agent_type], self.initializer.policy_function, agent_type)
                progress = self.controller.progress_inplace if self._inplace_progress else self.controller.progress
                self.state = progress(self.state, action_profile, self.initializer.transition_function)
                self.state_trajectory[-1].append(self._snapshot(self.state))

    # ------------------------------------------------------------------------- fused episode
    def _fused_plan(self):
        """Returns the modules of the SEIRM pattern if this config can run as one fused episode."""
        if getattr(self, "_plan_cache", None) is not None:
            plan = self._plan_cache
            return plan if plan and self._plan_still_valid(plan) else None
        self._plan_cache = False
        try:
            from agenttorch_b200.models.covid.substeps.new_transmission.action import MakeIsolationDecision
            from agenttorch_b200.models.covid.substeps.new_transmission.transition import NewTransmission
            from agenttorch_b200.models.covid.substeps.seirm_progression.transition import SEIRMProgression
        except ImportError:
            return None
        subs = self.config["substeps"]
        keys = list(subs.keys())
        if keys[:2] != ["0", "1"] or self.initializer.device.type != "cuda":
            return None
        tf, pf, of = self.initializer.transition_function, self.initializer.policy_function, self.initializer.observation_function
        if list(tf["0"].keys()) != ["new_transmission"] or list(tf["1"].keys()) != ["seirm_progression"]:
            return None
        tr, pg = tf["0"]["new_transmission"], tf["1"]["seirm_progression"]
        if type(tr) is not NewTransmission or type(pg) is not SEIRMProgression:
            return None
        agents = subs["0"]["active_agents"]
        if len(agents) != 1 or any(len(of[s][a]) for s in ("0", "1") for a in subs[s]["active_agents"]):
            return None
        pols = pf["0"][agents[0]]
        if list(pols.keys()) != ["make_isolation_decision"] or type(pols["make_isolation_decision"]) is not MakeIsolationDecision:
            return None
        if any(len(pf["1"][a]) for a in subs["1"]["active_agents"]):
            return None
        after_step = None
        if keys[2:]:
            # further substeps: only this package's quarantine / testing blocks, which ride along as one in-place launch each
            # after every step of the fused episode (models/covid/interventions.py)
            from agenttorch_b200.models.covid.interventions import plan_intervention_substeps

            after_step = plan_intervention_substeps(self, keys[2:])
            if after_step is None:
                return None
        self._plan_cache = {"transmission": tr, "progression": pg, "policy": pols["make_isolation_decision"], "after_step": after_step}
        return self._plan_cache if self._plan_still_valid(self._plan_cache) else None

    @staticmethod
    def _plan_still_valid(plan):
        """The parts of a cached plan a user can change between calls: an attached behaviour / injected draws."""
        if not plan["policy"].draws_uniform():
            return False
        after_step = plan.get("after_step")
        return after_step is None or getattr(after_step, "still_fusable", lambda: True)()

    def _step_fused(self, num_steps, plan):
        import re

        from agenttorch_b200 import ops
        from agenttorch_b200.core.helpers import get_by_path, set_by_path
        from agenttorch_b200.models.covid.context import as_kernel_f32, get_context

        tr, pg = plan["transmission"], plan["progression"]
        iv, pv = tr.input_variables, pg.input_variables
        state = self.state
        ctx = get_context(self.config, state, iv)

        split = plan.setdefault("_paths", {})             # "a/b/c" -> ["a", "b", "c"], once per plan

        def path(p):
            parts = split.get(p)
            if parts is None:
                parts = split[p] = re.split("/", p)
            return parts

        tensors = {}
        for name in ("disease_stage", "next_stage_time", "infected_time", "age"):
            cur = get_by_path(state, path(iv[name]))
            fixed = as_kernel_f32(cur)
            if fixed is not cur:
                set_by_path(state, path(iv[name]), fixed)
            if fixed.requires_grad or fixed.grad_fn is not None:
                fixed = fixed.detach()            # the fused path does not carry the stage tangent across calls
                set_by_path(state, path(iv[name]), fixed)
            tensors[name] = fixed
        di = get_by_path(state, path(iv["daily_infected"]))
        dd = get_by_path(state, path(pv["daily_deaths"]))
        if num_steps > di.numel():
            raise IndexError(f"num_steps={num_steps} exceeds the length of daily_infected ({di.numel()})")
        r_steps = tr.r_steps(num_steps)
        m_value = pg.m_tensor()
        want_grad = bool(r_steps.requires_grad or m_value.requires_grad) and torch.is_grad_enabled()
        tr.episode_index += 1
        uniform = None
        if tr.exposure_uniform is not None:
            uniform = torch.stack([as_kernel_f32(tr.exposure_uniform(t) if callable(tr.exposure_uniform)
                                                 else tr.exposure_uniform[t]).reshape(-1) for t in range(num_steps)])
        before = ctx.engine.launch_count
        make_hook = plan.get("after_step")
        hook = make_hook(state) if make_hook is not None else None
        ctx.engine.after_step = hook
        # SFSusceptibility / SFInfector given as differentiable tensors (learnable tables): their gradients come from the same
        # saved buffers as dL/dR (SeirmEngine.table_gradients)
        tables = []
        for name in ("SFSusceptibility", "SFInfector"):
            tv = get_by_path(state, path(iv.get(name, "environment/" + name)))
            tables.append(tv if torch.is_tensor(tv) and tv.requires_grad else None)
        want_grad = want_grad or (any(tv is not None for tv in tables) and torch.is_grad_enabled())
        try:
            new_di, new_dd = ops.seirm_episode_fn(
                ctx.handle, tensors["disease_stage"], tensors["next_stage_time"], tensors["infected_time"], tensors["age"],
                di.reshape(-1), dd.reshape(-1), r_steps, m_value, int(num_steps), tr.episode_seed(), "soft", 0.0,
                None, uniform, want_grad, self._use_cuda_graph and uniform is None, sfs=tables[0], sfi=tables[1])
        finally:
            ctx.engine.after_step = None
        # a lam_gamma_integrals overrun (the reference raises an IndexError) or an exchange timeout flagged on the device
        # surfaces here without a host synchronisation, i.e. on the call after the one that caused it
        ctx.engine.poll_errors()
        set_by_path(state, path(iv["daily_infected"]), new_di.view_as(di))
        set_by_path(state, path(pv["daily_deaths"]), new_dd.reshape(1, -1))       # (1,T) like the reference (Q6)
        state["current_step"] = num_steps - 1
        state["current_substep"] = "0"
        self.perf_stats["fused_episodes"] += 1
        self.perf_stats["kernel_launches"] += ctx.engine.launch_count - before + (hook.launches_per_step * num_steps if hook else 0)
        di_key, dd_key = path(iv["daily_infected"])[-1], path(pv["daily_deaths"])[-1]
        env = state["environment"]
        n_sub = len(self.config["substeps"])
        for t in range(num_steps):
            self.state_trajectory.append([])
            for s in range(n_sub):
                last = t == num_steps - 1 and s == n_sub - 1
                upto = {di_key: t + 1, dd_key: t + 1 if s >= 1 else t}          # deaths are written by substep '1'
                self.state_trajectory[-1].append({
                    "current_step": t, "current_substep": str((s + 1) % n_sub),
                    "environment": dict(env) if last else _LazyEnvironment(env, upto)})

    # ------------------------------------------------------------------------------- parameters
    def get_performance_stats(self):
        return {**self.perf_stats, "device": str(self.initializer.device)}

    def _set_parameters(self, params_dict):
        for name, value in params_dict.items():
            accessor = self._map_and_replace_tensor(name)
            current = accessor(self)
            if torch.is_tensor(value) and torch.is_tensor(current) and value.device != current.device:
                value = value.to(current.device)
            accessor(self, value)

    def _map_and_replace_tensor(self, input_string):
        """'initializer.transition_function.0.new_transmission.learnable_args.R2' -> attribute
        `calibrate_R2` of that substep module (runner.py:547-574)."""
        parts = input_string.split(".")
        function, index, sub_func, var_name = parts[1], parts[2], parts[3], parts[5]

        def getter_and_setter(runner, new_value=None):
            module = getattr(getattr(runner.initializer, function)[str(index)], sub_func)
            attr = "calibrate_" + var_name
            current = getattr(module, attr)
            if new_value is not None:
                assert new_value.requires_grad == current.requires_grad
                setattr(module, attr, new_value)
                return getattr(module, attr)
            return current

        return getter_and_setter

    def step_from_params(self, num_steps=None, params=None):
        if params is None:
            print(" missing parameters!!! ")
            return
        self._set_parameters(params)
        self.step(num_steps)

    def forward(self):
        for _ in range(self.config["simulation_metadata"]["num_episodes"]):
            self.step(self.config["simulation_metadata"]["num_steps_per_episode"])
