"""`vectorized` marker, `is_vectorized`, `VectorizedRunner` (reference: agent_torch/core/vectorization.py:14-57,
core/vectorized_runner.py:18-164) and the entry point `launch_distributed_simulation` (core/distributed_runner.py:285-316).

In the reference a VectorizedRunner looks up the `_vectorized` flag of every substep and then calls the same function either
way (vectorized_runner.py:67-164), so the class below only keeps the bookkeeping users can observe:
`is_vectorized` and the `vectorized_functions` table.  Every substep of this package is a batched CUDA kernel over all agents."""
from functools import wraps

from .runner import Runner


def vectorized(fn=None, *, in_dims=0, out_dims=0):
    """Mark a function or method as written for whole agent batches (`_vectorized`, `_vmap_in_dims`, `_vmap_out_dims`)."""

    def decorator(func):
        @wraps(func)
        def wrapped(*args, **kwargs):
            return func(*args, **kwargs)

        wrapped._vectorized = True
        wrapped._vmap_in_dims = in_dims
        wrapped._vmap_out_dims = out_dims
        wrapped._original_func = func
        return wrapped

    return decorator if fn is None else decorator(fn)


def is_vectorized(fn):
    return bool(getattr(fn, "_vectorized", False))


class VectorizedRunner(Runner):
    def __init__(self, config, registry) -> None:
        super().__init__(config, registry)
        self.is_vectorized = True
        self.vectorized_functions = {"observation": {}, "policy": {}, "transition": {}}

    def init(self):
        super().init()
        self._update_vectorized_functions()

    def _update_vectorized_functions(self):
        """Which substep modules declare themselves vectorized: the flag on the class, its forward, or `fusable` (the CUDA
        substeps of this package)."""
        init = self.initializer
        tables = {"observation": init.observation_function, "policy": init.policy_function}
        out = {"observation": {}, "policy": {}, "transition": {}}
        for kind, table in tables.items():
            for substep, per_agent in table.items():
                for agent_type, mods in per_agent.items():
                    for name, mod in mods.items():
                        out[kind][f"{substep}/{agent_type}/{name}"] = self._flag(mod)
        for substep, mods in init.transition_function.items():
            for name, mod in mods.items():
                out["transition"][f"{substep}/{name}"] = self._flag(mod)
        self.vectorized_functions = out

    @staticmethod
    def _flag(module):
        return bool(is_vectorized(module) or is_vectorized(getattr(module, "forward", None)) or getattr(module, "fusable", False))


def launch_distributed_simulation(config, registry, world_size=None, num_steps=None):
    """Single-process call of the reference's launcher.  With one process the episode runs here and the final state is
    returned; several GPUs need one process per GPU started by `torchrun` (this package never spawns processes: the
    reference re-spawns and re-initialises on every step, core/environment.py:31-47) — then the same call, made on every
    rank, runs the sharded episode and returns the state on rank 0."""
    import torch
    import torch.distributed as dist

    runner = Runner(config, registry)
    launched = dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1
    if not launched:
        if world_size not in (None, 0, 1) and torch.cuda.device_count() > 1:
            raise RuntimeError(f"world_size={world_size}: start one process per GPU, e.g. `torchrun --nproc-per-node {world_size} "
                               "script.py`, and call launch_distributed_simulation on every rank")
        runner.init()
        runner.step(num_steps)
        return runner.state
    from agenttorch_b200.distributed import DistributedRunner

    drunner = DistributedRunner(runner, world_size=world_size)
    drunner.init()
    drunner.step(num_steps)
    return drunner.state if dist.get_rank() == 0 else None
