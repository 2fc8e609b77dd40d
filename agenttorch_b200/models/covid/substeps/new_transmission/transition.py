"""NewTransmission — substep '0' transition, B200 path.

Drop-in for the reference class of the same name (models/covid/substeps/new_transmission/
transition.py:11-252): same constructor, same `forward(state, action) -> {output_variables[i]: tensor}`,
same `calibrate_R2` / `learnable_args["R2"]` parameter handling (:139-143).  The per-edge PyG gather,
the scatter-add, the Bernoulli draw and the stage/timer updates all run in one CUDA launch
(`agenttorch_b200::seirm_transmission`); autograd is registered on the op.

Randomness: the reference draws `torch.bernoulli(p)` from the global generator.  Here the exposure
uniform comes from the counter hash keyed by (seed, episode, step, agent), or from `exposure_uniform`
when a test injects it (a callable t -> (N,) tensor, the analogue of overriding `st_bernoulli`)."""
import re

import torch

from agenttorch_b200 import ops
from agenttorch_b200.core.helpers import get_by_path
from agenttorch_b200.core.substep import SubstepTransitionMessagePassing
from agenttorch_b200.models.covid.context import as_kernel_f32, get_context


class _StepR(torch.autograd.Function):
    """r_steps[t] = (R2 * one_hot[week(t)]).sum() (transition.py:136-143), evaluated once per distinct week and gathered per step.
    All weeks in one broadcast product and one row-wise sum (bit-identical to the reference's per-step expression on this device:
    tools/probe_host.py compares the two on random R2); the Jacobian d r_steps / d R2 is a constant `jac` of shape
    (T, *R2.shape), so backward is one product and one sum."""

    @staticmethod
    def forward(ctx, r, eye_used, pos, jac):
        per_week = (r.detach().unsqueeze(0) * eye_used).flatten(1).sum(1)
        ctx.jac = jac
        return per_week.index_select(0, pos)

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, g):
        jac = ctx.jac
        if jac is None:
            return None, None, None, None
        return (jac * g.reshape(-1, *([1] * (jac.dim() - 1)))).sum(0), None, None, None


class NewTransmission(SubstepTransitionMessagePassing):
    fusable = True

    def __init__(self, config, input_variables, output_variables, arguments):
        super().__init__(config, input_variables, output_variables, arguments)
        meta = self.config["simulation_metadata"]
        self.device = torch.device(meta["device"])
        self.num_timesteps = meta["num_steps_per_episode"]
        self.num_weeks = meta["NUM_WEEKS"]
        self.EXPOSED_TO_INFECTED_TIME = meta["EXPOSED_TO_INFECTED_TIME"]
        self.mode = meta["EXECUTION_MODE"]
        self.calibration_mode = meta["calibration"]
        self._week_one_hot = torch.eye(self.num_weeks, device=self.device)
        seed = meta.get("noise_seed")
        self.noise_seed = int(seed) if seed is not None else int(torch.randint(0, 2**62, (1,)).item())
        self.episode_index = 0
        self.advance_seed = True          # a fresh noise stream per episode, like the reference's global RNG
        self.exposure_uniform = None
        self._r_index = {}

    def r_tensor(self):
        return self.calibrate_R2 if self.calibration_mode else self.learnable_args["R2"]

    def effective_r(self, t):
        """R of step t exactly as the reference forms it (:136-143): with the YAML's (NUM_WEEKS,1) R2 the
        product broadcasts and `.sum()` adds every week (Q2); with a (NUM_WEEKS,) tensor it is R2[week]."""
        return (self.r_tensor() * self._week_one_hot[int(t / 7)]).sum()

    def r_steps(self, num_steps):
        """(num_steps,) per-step R with the values of `effective_r(t)`; 3 launches forward, 2 backward (autograd through 21 stacked
        scalars took ~50).  Index tensors and the constant Jacobian are cached per (num_steps, R2.shape)."""
        r = self.r_tensor()
        key = (int(num_steps), tuple(r.shape), r.device)
        cached = self._r_index.get(key)
        if cached is None:
            eye = self._week_one_hot.to(r.device)
            weeks = [int(t / 7) for t in range(int(num_steps))]
            used = sorted(set(weeks))
            pos = torch.tensor([used.index(w) for w in weeks], dtype=torch.long, device=r.device)
            # one_hot rows of the weeks in use, shaped to broadcast against R2 exactly as `R2 * one_hot[w]` does
            eye_used = eye[used].reshape(len(used), *([1] * max(r.dim() - 1, 0)), eye.shape[1]).contiguous()
            full = torch.broadcast_shapes(tuple(r.shape), (eye.shape[1],))
            jac = torch.stack([torch.broadcast_to(eye[w], full).sum_to_size(r.shape) for w in weeks]).contiguous()
            cached = self._r_index[key] = (eye_used, pos, jac)
        eye_used, pos, jac = cached
        return _StepR.apply(r, eye_used, pos, jac if r.requires_grad else None)

    def episode_seed(self):
        """64-bit seed of this episode's noise stream, as a signed int64 (custom-op integer arguments are int64)."""
        v = self.noise_seed if not self.advance_seed else self.noise_seed + 0x9E3779B97F4A7C15 * self.episode_index
        v &= (1 << 64) - 1
        return v - (1 << 64) if v >= (1 << 63) else v

    def forward(self, state, action=None):
        iv = self.input_variables
        t = int(state["current_step"])
        if t == 0:
            self.episode_index += 1
        ctx = get_context(self.config, state, iv)

        def var(name):
            return get_by_path(state, re.split("/", iv[name]))

        for name in ("SFSusceptibility", "SFInfector"):
            tv = get_by_path(state, re.split("/", iv.get(name, "environment/" + name)))
            if torch.is_tensor(tv) and tv.requires_grad and torch.is_grad_enabled():
                raise NotImplementedError(f"gradients with respect to {name} are computed by the fused episode path (Runner.step with "
                                          "cuda_params.fused_episode), not by the per-substep operators")
        stage = as_kernel_f32(var("disease_stage"))
        nst = as_kernel_f32(var("next_stage_time"))
        itime = as_kernel_f32(var("infected_time"))
        age = as_kernel_f32(var("age"))
        daily_infected = var("daily_infected")
        will_isolate = as_kernel_f32(action["citizens"]["isolation_decision"])
        uniform = None
        if self.exposure_uniform is not None:
            uniform = as_kernel_f32(self.exposure_uniform(t) if callable(self.exposure_uniform) else self.exposure_uniform[t])
        r = self.effective_r(t)
        new_stage, new_nst, new_itime, new_di, _ = ops.seirm_transmission(
            ctx.handle, stage, nst, itime, age, daily_infected, r.reshape(1), t, self.episode_seed(), "tensor", 0.0,
            will_isolate, uniform, bool(r.requires_grad))
        ov = self.output_variables
        return {ov[0]: new_stage, ov[1]: new_nst, ov[2]: new_itime, ov[3]: new_di}
