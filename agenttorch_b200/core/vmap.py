"""`@vmap` class decorator and `sample_grid` (reference: agent_torch/core/helpers/vmap.py:33-250).

`@vmap(agent_args, shared_args, outputs)` wraps a substep whose forward(state, action) is written
for ONE agent: per-agent fields are taken from state["agents"][<first agent type>], shared fields from
state["environment"] (falling back to agent state), learnable/fixed arguments are added to the
per-agent dict, and torch.vmap(in_dims=0, randomness="different") runs the batch."""
import torch
import torch.nn.functional as F


def sample_grid(grid, y, x):
    """Bilinear read of grid[y, x] at fractional coordinates, differentiable, vmap-safe."""
    h, w = grid.shape[-2], grid.shape[-1]
    yn = 2.0 * y / max(h - 1, 1) - 1.0
    xn = 2.0 * x / max(w - 1, 1) - 1.0
    pt = torch.stack([xn, yn]).view(1, 1, 1, 2).to(grid.dtype)
    return F.grid_sample(grid.reshape(1, 1, h, w), pt, mode="bilinear", padding_mode="border", align_corners=True).squeeze()


def vmap(agent_args, shared_args=None, outputs=None, compile=False):
    shared_args = list(shared_args or [])
    outputs = list(outputs or [])

    def decorator(cls):
        cls._vmap_config = {"agent_args": list(agent_args), "shared_args": shared_args, "outputs": outputs, "compile": compile}
        single_agent_forward = cls.forward

        def batched_forward(self, state, action=None):
            agents = state["agents"]
            agent_type = next(iter(agents))
            props = agents[agent_type]
            per_agent = {k: props[k] for k in agent_args}
            shared = {}
            for k in shared_args:
                env = state.get("environment") or {}
                shared[k] = env[k] if k in env else props[k]
            extra = {}
            for src in (getattr(self, "learnable_args", None) or {}, getattr(self, "fixed_args", None) or {}):
                for k, v in src.items():
                    extra[k] = v

            def one(agent_fields):
                view = dict(shared)
                view.update(extra)
                view.update(agent_fields)
                res = single_agent_forward(self, view, action)
                return tuple(res[o] for o in outputs)

            fn = torch.vmap(one, in_dims=({k: 0 for k in per_agent},), randomness="different")
            if compile:
                fn = torch.compile(fn)
            res = fn(per_agent)
            names = self.output_variables if getattr(self, "output_variables", None) else outputs
            return {names[i] if i < len(names) else outputs[i]: res[i] for i in range(len(outputs))}

        cls.forward = batched_forward
        cls._single_agent_forward = single_agent_forward
        return cls

    return decorator
