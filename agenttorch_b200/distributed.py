"""DistributedRunner — multi-GPU SEIRM episodes behind the Runner API.

What the reference offers here (agent_torch/core/distributed_runner.py:23-316, `envs.create(distributed=True,
world_size=...)`, core/environment.py:85-123): every rank initialises the full state, keeps the contiguous slice
`[start, end)` of every agent tensor, steps, and rank 0 all-gathers the final agent state; interactions that cross
a partition boundary are a stub (`_handle_cross_gpu_interactions`, :218-228), so network substeps cannot run.

Here the same surface drives the sharded engine: one process per GPU (started by `torchrun`, never spawned from
inside), contiguous ownership aligned to 1024 agents, rows of the contact graph sharded with GLOBAL source ids, and the
real per-step halo exchange of the infectiousness codes (symmetric-memory barrier + P2P pull, or NCCL all-gather).
`step(n)` advances the local slices in place and leaves the all-reduced daily series (with autograd to R2 / M) in
`state["environment"]`; `gather_final_state()` reassembles the agent tensors on every rank.
"""
import re

import torch
import torch.distributed as dist

from .core.helpers import get_by_path
from .engine import NoiseSpec, SeirmEngine, SeirmState
from .graph import build_contact_graph, shard_bounds, shard_code_len, shard_graph


def _require_cuda(is_cuda):
    """No CPU fallback.  (One function so that the host-emulation tests can lift the check.)"""
    if not is_cuda:
        raise RuntimeError("DistributedRunner needs simulation_metadata.device: cuda (no CPU fallback)")


class _EpisodeFn(torch.autograd.Function):
    """Differentiable multi-GPU episode: forward advances the local shard, backward runs the fused backward kernel and
    all-reduces the parameter gradients."""

    @staticmethod
    def forward(ctx, runner, r_steps, m_value, num_steps, want_grad):
        eng, st = runner.engine, runner.local_state
        noise = NoiseSpec(seed=runner.episode_seed(), isolation="soft")
        factory = getattr(runner, "_after_step_factory", None)
        eng.after_step = factory(runner.state) if factory is not None else None      # episodes with a hook are issued eagerly
        try:
            # the local series are zeroed inside the episode's (captured) launch sequence
            sv = eng.run_episode(st, r_steps, m_value, num_steps, noise, want_grad=want_grad, cuda_graph=eng.code_hdl is not None,
                                 zero_series=runner._series)
        finally:
            eng.after_step = None
        if eng.pg is not None:
            from .parallel import reduce_daily_series

            di, dd = reduce_daily_series(st, eng.pg)
        else:
            di, dd = st.daily_infected.clone(), st.daily_deaths.clone()
        ctx.runner, ctx.sv, ctx.num_steps, ctx.want_grad = runner, sv, num_steps, want_grad
        ctx.save_for_backward(m_value)
        ctx.r_shape, ctx.m_shape = r_steps.shape, m_value.shape
        return di, dd

    @staticmethod
    def backward(ctx, g_di, g_dd):
        (m_value,) = ctx.saved_tensors
        if not ctx.want_grad:
            return None, None, None, None, None
        T = ctx.num_steps
        gr, gm = ctx.runner.engine.episode_backward(ctx.sv, T, g_di, g_dd, m_value)     # (the engine flattens / trims if it has to)
        g_m = gm.reshape(ctx.m_shape) if m_value.numel() == 1 else gm.expand(m_value.numel()).reshape(ctx.m_shape).clone()
        return None, gr.reshape(ctx.r_shape), g_m, None, None


class DistributedRunner:
    def __init__(self, runner, world_size=None, distributed_config=None, process_group=None):
        self.runner = runner
        self.config = runner.config
        self.registry = runner.registry
        self.initializer = runner.initializer
        self.world_size = world_size
        self.distributed_config = distributed_config or {"strategy": "data_parallel", "sync_frequency": 1}
        self.pg = process_group
        self.engine = None
        self.local_state = None
        self._after_step_factory = None
        self.state = None
        self.state_trajectory = []
        self._episode = 0

    # ------------------------------------------------------------------------------------------ setup
    def _group(self):
        if self.pg is not None:
            return self.pg
        if dist.is_available() and dist.is_initialized():
            return dist.group.WORLD if dist.get_world_size() > 1 else None
        return None

    def init(self):
        """Full initialisation on every rank (as the reference does), then slice to the owned agents."""
        self.runner.init()
        full = self.runner.state
        self.pg = self._group()
        world = dist.get_world_size(self.pg) if self.pg is not None else 1
        rank = dist.get_rank(self.pg) if self.pg is not None else 0
        if self.world_size not in (None, world):
            raise RuntimeError(f"world_size={self.world_size} requested but the process group has {world} ranks "
                               "(launch one process per GPU with torchrun)")
        self.world_size, self.rank = world, rank
        plan = self.runner._fused_plan()
        if plan is None:
            raise NotImplementedError("distributed execution is implemented for the fused SEIRM path only (transmission + "
                                      "progression, optionally followed by this package's quarantine / testing substeps)")
        # quarantine / testing substeps ride along: per-agent state, sliced like every agent tensor below, draws addressed by
        # the GLOBAL agent id (id_offset = first owned agent), one in-place launch per block after every step
        self._after_step_factory = plan.get("after_step")
        self.tr, self.pg_mod = plan["transmission"], plan["progression"]
        iv = self.tr.input_variables
        n = int(self.config["simulation_metadata"]["num_agents"])
        T = int(self.config["simulation_metadata"]["num_steps_per_episode"])
        self.num_agents, self.T = n, T
        self.bounds = shard_bounds(n, world)
        r0, r1 = self.bounds[rank]
        self.agent_partition_map = {"start_idx": r0, "end_idx": r1, "local_size": r1 - r0, "global_size": n}

        def var(name):
            return get_by_path(full, re.split("/", iv[name]))

        edge_index, edge_attr = var("adjacency_matrix")
        _require_cuda(edge_index.is_cuda)
        from .graph import ContactGraph
        from .models.covid.context import spec_from_state

        net = get_by_path(full, re.split("/", iv["adjacency_matrix"])[:-1])
        # distributed_config["partition"] = {"parts": P, "rounds": R} (or True: P = world size): re-order the population so that
        # the ranks' contiguous ranges are communities of the contact graph (partition.py) before anything is sliced.  The agents
        # get NEW ids (the noise streams follow them): the same `parts` gives the same population at every world size;
        # gather_final_state() hands the agent tensors back in the original order.
        self._order = None
        part_cfg = (self.distributed_config or {}).get("partition")
        if part_cfg:
            from . import partition

            part_cfg = part_cfg if isinstance(part_cfg, dict) else {}
            labels = partition.label_propagation_partition(edge_index, n, int(part_cfg.get("parts", world)),
                                                           rounds=int(part_cfg.get("rounds", 10)))
            agent_type = self.config["substeps"]["0"]["active_agents"][0]
            per_agent = {k: v for k, v in full["agents"][agent_type].items()
                         if torch.is_tensor(v) and v.dim() > 0 and v.shape[0] == n}
            # (the reference keeps one per-agent tensor in the environment: mean_interactions, config.yaml:219)
            per_env = {k: v for k, v in full["environment"].items() if torch.is_tensor(v) and v.dim() > 0 and v.shape[0] == n}
            new_edges, moved, order = partition.reorder_population(edge_index, {**{("a", k): v for k, v in per_agent.items()},
                                                                                **{("e", k): v for k, v in per_env.items()}}, labels)
            for (where, k), v in moved.items():
                (full["agents"][agent_type] if where == "a" else full["environment"])[k] = v
            edge_index = new_edges
            net["adjacency_matrix"] = (edge_index, edge_attr)                 # (edge order unchanged: edge_attr still applies)
            net.pop("graph", None)                                            # a prebuilt CSR describes the old numbering
            self._order = order
            self.partition_report = {"parts": int(part_cfg.get("parts", world)),
                                     "edge_cut": partition.edge_cut_fraction(edge_index[0].long(), edge_index[1].long(),
                                                                             labels[order])}
        prebuilt = net.get("graph") if isinstance(net, dict) else None
        if isinstance(prebuilt, ContactGraph) and prebuilt.num_edges == edge_index.shape[1] and prebuilt.row_offset == 0:
            graph_full = prebuilt.to(edge_index.device)          # the network helper already produced the dst-sorted CSR
        else:
            graph_full = build_contact_graph(edge_index, n, edge_attr=edge_attr, mean_interactions=var("mean_interactions"),
                                             device=edge_index.device)
        shard = shard_graph(graph_full, r0, r1)
        del graph_full
        self.engine = SeirmEngine(shard, spec_from_state(self.config, full, iv), process_group=self.pg,
                                  code_len=shard_code_len(n, world) if world > 1 else None)
        f32 = lambda t: t.reshape(-1)[r0:r1].to(torch.float32).clone().contiguous()          # noqa: E731
        dev = edge_index.device
        self._series = torch.zeros((2, T), dtype=torch.float32, device=dev)      # this rank's share of the two daily series
        self.local_state = SeirmState(f32(var("disease_stage")), f32(var("next_stage_time")), f32(var("infected_time")),
                                      f32(var("age")), self._series[0], self._series[1])
        # the public state holds the LOCAL slices of the agent tensors (reference: _slice_agent_tensors, :126-144)
        cit = full["agents"][self.config["substeps"]["0"]["active_agents"][0]]
        for k, v in list(cit.items()):
            if torch.is_tensor(v) and v.dim() > 0 and v.shape[0] == n:
                cit[k] = v[r0:r1]
        cit["disease_stage"] = self.local_state.stage.view(-1, 1)
        cit["next_stage_time"] = self.local_state.next_stage_time.view(-1, 1)
        cit["infected_time"] = self.local_state.infected_time.view(-1, 1)
        if self._after_step_factory is not None:
            for _, mod in self.runner._noise_modules():
                if hasattr(mod, "id_offset"):
                    mod.id_offset = r0
        self.config["simulation_metadata"]["local_num_agents"] = r1 - r0
        self.state = full
        self.state_trajectory = [[{"current_step": 0, "current_substep": "0", "environment": dict(full["environment"])}]]

    def reset(self):
        self.init()

    def local_snapshot(self):
        """The owned slices of the agent state and the daily series, as `restore_state` takes them back."""
        st = self.local_state
        agent = self.config["substeps"]["0"]["active_agents"][0]
        return {"agents": {agent: {"disease_stage": st.stage.view(-1, 1).clone(), "next_stage_time": st.next_stage_time.view(-1, 1).clone(),
                                   "infected_time": st.infected_time.view(-1, 1).clone()}},
                "environment": {"daily_infected": self.state["environment"]["daily_infected"].detach().clone(),
                                "daily_deaths": self.state["environment"]["daily_deaths"].detach().clone()}}

    def restore_state(self, snapshot, non_blocking=True):
        """Copy a saved state of THIS rank's agents (nested dict of tensors or a core.snapshot.PackedSnapshot, host or device)
        into the live local tensors — the sharded counterpart of Runner.restore_state; the engine bindings and captured CUDA
        graphs stay valid."""
        from .core.snapshot import PackedSnapshot

        env = self.state["environment"]
        for k in ("daily_infected", "daily_deaths"):                  # step() replaced them by autograd outputs
            if env[k].requires_grad or env[k].grad_fn is not None:
                env[k] = env[k].detach().clone()
        if isinstance(snapshot, PackedSnapshot):
            snapshot.restore_into(self.state, non_blocking=non_blocking)
        else:
            def walk(dst, src):
                for k, v in src.items():
                    if isinstance(v, dict):
                        walk(dst[k], v)
                    elif torch.is_tensor(v) and torch.is_tensor(dst.get(k)):
                        cur = dst[k]
                        if v.dtype != cur.dtype and v.device != cur.device:
                            cur.copy_(v.to(cur.device, non_blocking=non_blocking).reshape(cur.shape))
                        else:
                            cur.copy_(v.reshape(cur.shape), non_blocking=non_blocking)
            walk(self.state, snapshot)
        self.state["current_step"] = 0

    def episode_seed(self):
        self.tr.episode_index = self._episode
        return self.tr.episode_seed()

    # ------------------------------------------------------------------------------------------- step
    def step(self, num_steps=None):
        if self.engine is None:
            raise RuntimeError("Must call init() before step()")
        T = num_steps or self.T
        self._episode += 1
        r_steps = self.tr.r_steps(T)
        m_value = self.pg_mod.m_tensor()
        want_grad = bool(r_steps.requires_grad or m_value.requires_grad) and torch.is_grad_enabled()
        env = self.state["environment"]
        di, dd = _EpisodeFn.apply(self, r_steps, m_value, T, want_grad)
        self.engine.poll_errors()                 # device-side error flag of an earlier episode (no host synchronisation)
        env["daily_infected"] = env["daily_infected"].detach().reshape(-1) + di
        env["daily_deaths"] = (env["daily_deaths"].detach().reshape(-1) + dd).reshape(1, -1)         # (1,T), Q6
        self.state["current_step"] = T - 1
        self.state_trajectory.append([{"current_step": T - 1, "current_substep": "0", "environment": dict(env)}])
        return self.state

    def forward(self):
        meta = self.config["simulation_metadata"]
        for episode in range(meta["num_episodes"]):
            if episode > 0:
                self.reset()
            self.step(meta["num_steps_per_episode"])

    def _set_parameters(self, params):
        self.runner._set_parameters(params)

    # ---------------------------------------------------------------------------------------- gather
    def gather_agent_states(self, local_tensor):
        """All-gather a local agent tensor into the full population (reference: gather_agent_states, :157-166)."""
        if self.pg is None:
            return local_tensor
        per = self.bounds[0][1] - self.bounds[0][0]
        pad = torch.zeros((per,) + tuple(local_tensor.shape[1:]), dtype=local_tensor.dtype, device=local_tensor.device)
        pad[: local_tensor.shape[0]] = local_tensor
        out = torch.empty((per * self.world_size,) + tuple(local_tensor.shape[1:]), dtype=local_tensor.dtype, device=local_tensor.device)
        dist.all_gather_into_tensor(out, pad, group=self.pg)
        return out[: self.num_agents]

    def gather_final_state(self):
        self.engine.check_errors()                # e.g. a halo-exchange rendezvous that timed out (device-side flag)
        cit = self.state["agents"][self.config["substeps"]["0"]["active_agents"][0]]
        out = {k: self.gather_agent_states(v) for k, v in cit.items()
               if torch.is_tensor(v) and v.dim() > 0 and v.shape[0] == self.agent_partition_map["local_size"]}
        if getattr(self, "_order", None) is not None:                         # back to the population's original agent order
            for k, v in out.items():
                orig = torch.empty_like(v)
                orig[self._order.to(v.device)] = v
                out[k] = orig
        return out
