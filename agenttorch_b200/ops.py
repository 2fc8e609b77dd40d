"""torch.library custom ops (namespace `agenttorch_b200`) over the C ABI, with registered autograd.

Three operators replace the reference's ATen op chains and their autograd graph:
  seirm_transmission  — NewTransmission.forward       (new_transmission/transition.py:131-252)
  seirm_progression   — SEIRMProgression.forward      (seirm_progression/transition.py:90-121)
  seirm_episode       — T x (transmission + progression) fused, state advanced in place (declared in the schema), with a fused
                        O(N) backward (at_seirm_episode_backward) instead of a T-step autograd tape.
An engine (graph + buffers) is addressed by an integer handle because custom ops only take tensors
and scalars.  Every op launches CUDA kernels from libat_seirm.so; there is no CPU implementation —
calling them with CPU tensors raises.
"""
from typing import Optional, Tuple

import torch
from torch import Tensor

from . import _native as nat
from .engine import NoiseSpec, SeirmEngine, SeirmState

_ENGINES = {}
_NEXT_HANDLE = [1]


def register_engine(engine: SeirmEngine) -> int:
    h = _NEXT_HANDLE[0]
    _NEXT_HANDLE[0] += 1
    _ENGINES[h] = engine
    return h


def release_engine(handle: int):
    _ENGINES.pop(handle, None)


def get_engine(handle: int) -> SeirmEngine:
    return _ENGINES[handle]


def _noise(seed, iso_mode, p_isolate, isolation, uniform):
    return NoiseSpec(seed=seed, isolation="tensor" if isolation is not None else iso_mode, p_isolate=p_isolate,
                     uniform="tensor" if uniform is not None else "hash", isolation_tensor=isolation, uniform_tensor=uniform)


def _flat_state(stage, nst, itime, age, di, dd):
    return SeirmState(stage.view(-1), nst.view(-1), itime.view(-1), age.view(-1), di.view(-1), dd.view(-1))


# ------------------------------------------------------------------------------------ transmission
@torch.library.custom_op("agenttorch_b200::seirm_transmission", mutates_args=())
def seirm_transmission(handle: int, stage: Tensor, next_stage_time: Tensor, infected_time: Tensor, age: Tensor,
                       daily_infected: Tensor, r_value: Tensor, t: int, seed: int, iso_mode: str, p_isolate: float,
                       isolation: Optional[Tensor], uniform: Optional[Tensor], want_grad: bool
                       ) -> Tuple[Tensor, Tensor, Tensor, Tensor, Tensor]:
    """Substep '0'.  Returns (stage, next_stage_time, infected_time, daily_infected, dne_dr) as NEW tensors."""
    eng = get_engine(handle)
    st = _flat_state(stage.detach().clone(), next_stage_time.detach().clone(), infected_time.detach().clone(), age,
                     daily_infected.detach().clone(), daily_infected.new_zeros(daily_infected.numel()))
    dne = torch.zeros(st.stage.numel(), dtype=torch.float32, device=stage.device) if want_grad else None
    eng.encode(st, t)   # the state may have been edited between calls: re-derive the infectiousness of step t
    noise = _noise(seed, iso_mode, p_isolate, None if isolation is None else isolation.reshape(-1).contiguous(),
                   None if uniform is None else uniform.reshape(-1).contiguous())
    eng.launch(st, flags=nat.DO_TRANSMISSION, t=t, noise=noise, isolation=noise.isolation_tensor,
               uniform=noise.uniform_tensor, r_value=r_value.detach().reshape(-1).to(torch.float32).contiguous(),
               r_index=0, dne_dr=dne)
    if dne is None:
        dne = torch.zeros(0, dtype=torch.float32, device=stage.device)
    return (st.stage.view_as(stage), st.next_stage_time.view_as(next_stage_time), st.infected_time.view_as(infected_time),
            st.daily_infected.view_as(daily_infected), dne)


@seirm_transmission.register_fake
def _(handle, stage, next_stage_time, infected_time, age, daily_infected, r_value, t, seed, iso_mode, p_isolate,
      isolation, uniform, want_grad):
    n = stage.numel() if want_grad else 0
    return (torch.empty_like(stage), torch.empty_like(next_stage_time), torch.empty_like(infected_time),
            torch.empty_like(daily_infected), stage.new_empty(n, dtype=torch.float32))


def _transmission_setup(ctx, inputs, output):
    ctx.t = inputs[7]
    ctx.r_shape = inputs[6].shape
    ctx.mark_non_differentiable(output[1], output[2], output[4])
    ctx.save_for_backward(output[4])


def _transmission_backward(ctx, g_stage, g_nst, g_itime, g_di, g_dne):
    (dne,) = ctx.saved_tensors
    g_r = None
    if dne.numel():
        # d ne_i / dR feeds stage_i (+= ne_i) and daily_infected[t] (+= sum ne)   (SURVEY.md A.1 reverse form)
        w = torch.zeros_like(dne) if g_stage is None else g_stage.reshape(-1).to(torch.float32)
        if g_di is not None:
            w = w + g_di.reshape(-1)[ctx.t].to(torch.float32)
        g_r = (dne * w).sum().reshape(ctx.r_shape)
    return (None, g_stage, None, None, None, g_di, g_r, None, None, None, None, None, None, None)


seirm_transmission.register_autograd(_transmission_backward, setup_context=_transmission_setup)


# ------------------------------------------------------------------------------------- progression
@torch.library.custom_op("agenttorch_b200::seirm_progression", mutates_args=())
def seirm_progression(handle: int, stage: Tensor, next_stage_time: Tensor, daily_deaths: Tensor, m_value: Tensor,
                      t: int) -> Tuple[Tensor, Tensor, Tensor, Tensor, Tensor]:
    """Substep '1'.  Returns (stage, next_stage_time, daily_deaths, recovered_mask u8, death_count (1,))."""
    eng = get_engine(handle)
    dev = stage.device
    nst = next_stage_time.detach().clone()
    st = SeirmState(stage.detach().clone().view(-1), nst.view(-1), nst.view(-1), nst.view(-1),
                    daily_deaths.new_zeros(daily_deaths.numel()), daily_deaths.detach().clone().view(-1))
    st.age = torch.zeros(1, dtype=torch.uint8, device=dev)          # not read by the progression pass
    rstep = torch.full((st.stage.numel(),), 255, dtype=torch.uint8, device=dev)
    dcount = torch.zeros(max(daily_deaths.numel(), t + 1), dtype=torch.float32, device=dev)
    # DO_INIT: the E / I row bitmaps are derived from THIS stage tensor (the engine's cached bitmaps describe whatever state
    # its previous launch saw, which need not be this one: a fresh engine, a state edited between substeps, two states)
    eng.launch(st, flags=nat.DO_PROGRESSION | nat.DO_INIT, t=t, m_value=m_value.detach().reshape(-1).to(torch.float32).contiguous(),
               death_count=dcount, recovered_step=rstep)
    return (st.stage.view_as(stage), st.next_stage_time.view_as(next_stage_time), st.daily_deaths.view_as(daily_deaths),
            (rstep == t).to(torch.uint8), dcount[t:t + 1].clone())


@seirm_progression.register_fake
def _(handle, stage, next_stage_time, daily_deaths, m_value, t):
    return (torch.empty_like(stage), torch.empty_like(next_stage_time), torch.empty_like(daily_deaths),
            stage.new_empty(stage.numel(), dtype=torch.uint8), stage.new_empty(1, dtype=torch.float32))


def _progression_setup(ctx, inputs, output):
    ctx.t = inputs[5]
    ctx.m_shape = inputs[4].shape
    ctx.stage_shape = inputs[1].shape
    ctx.mark_non_differentiable(output[1], output[3], output[4])
    ctx.save_for_backward(output[3], output[4], inputs[4])


def _progression_backward(ctx, g_stage, g_nst, g_dd, g_mask, g_cnt):
    rd, dcount, m = ctx.saved_tensors
    g_stage_in, g_m = g_stage, None
    if g_dd is not None:
        gt = g_dd.reshape(-1)[ctx.t]
        # daily_deaths[t] += M * sum(stage * rd / 2): d/dstage_i = M*rd_i/2, d/dM = sum(stage*rd/2)
        add = (gt * m.reshape(-1)[0] * 0.5) * rd.to(gt.dtype).view(ctx.stage_shape)
        g_stage_in = add if g_stage is None else g_stage + add
        g_m = (gt * dcount.reshape(-1)[0]).reshape(1).expand(m.numel()).reshape(ctx.m_shape).clone()
    return (None, g_stage_in, None, g_dd, g_m, None)


seirm_progression.register_autograd(_progression_backward, setup_context=_progression_setup)


# ----------------------------------------------------------------------------------------- episode
# The episode operator MUTATES the agent-state tensors it is given (stage, timers) and says so in its schema, so the dispatcher
# (functionalisation, torch.compile) knows.  torch.library only registers autograd formulas for functional operators, so the
# differentiable entry point is an autograd.Function around the mutating operator (`seirm_episode_fn` below): its forward
# calls the operator, its backward the fused backward kernel — the state tensors are never differentiable inputs.
@torch.library.custom_op("agenttorch_b200::seirm_episode", mutates_args=("stage", "next_stage_time", "infected_time"))
def seirm_episode(handle: int, stage: Tensor, next_stage_time: Tensor, infected_time: Tensor, age: Tensor,
                  daily_infected: Tensor, daily_deaths: Tensor, r_steps: Tensor, m_value: Tensor,
                  num_steps: int, seed: int, iso_mode: str, p_isolate: float, isolation: Optional[Tensor],
                  uniform: Optional[Tensor], want_grad: bool, cuda_graph: bool, want_table_grad: bool) -> Tuple[Tensor, Tensor]:
    """Advance the agent state IN PLACE through `num_steps` steps.  Returns the new (daily_infected, daily_deaths); everything
    saved for backward (the (T,N) dne_dr, recovered_step, the (T,) sum_dne_dr and death_count) stays inside the engine until
    its next episode (`SeirmEngine.saved_buffers`)."""
    eng = get_engine(handle)
    # the kernels accumulate this episode's increments into an engine-owned (2, L) buffer that the episode's own launch sequence
    # zeroes: a captured episode is found again whatever address the caller's (freshly cloned, autograd-carrying) series tensors
    # have, and the host issues one launch for the two results.  The flattened state views are cached per set of state tensors
    # (every view is a dispatcher call, and the host - not the GPU - bounds a small shard's episode).
    n_di, n_dd = daily_infected.numel(), daily_deaths.numel()
    ck = (stage.data_ptr(), next_stage_time.data_ptr(), infected_time.data_ptr(), age.data_ptr(), stage.numel(), stage.dtype,
          next_stage_time.dtype, infected_time.dtype, age.dtype, n_di, n_dd)
    cached = getattr(eng, "_op_state", None)
    if cached is None or cached[0] != ck:
        bufs = torch.zeros((2, max(n_di, n_dd)), dtype=torch.float32, device=daily_infected.device)
        st = SeirmState(stage.view(-1), next_stage_time.view(-1), infected_time.view(-1), age.view(-1), bufs[0, :n_di], bufs[1, :n_dd])
        cached = eng._op_state = (ck, st, bufs)
    _, st, bufs = cached
    noise = _noise(seed, iso_mode, p_isolate, isolation, uniform)
    eng.want_table_grad = bool(want_table_grad)
    eng.run_episode(st, r_steps, m_value, num_steps, noise, want_grad=want_grad, cuda_graph=cuda_graph, zero_series=bufs)
    if (daily_infected.dtype == torch.float32 and daily_deaths.dtype == torch.float32 and daily_infected.dim() == 1
            and daily_deaths.dim() == 1 and daily_infected.is_contiguous() and daily_deaths.is_contiguous()):
        new_di, new_dd = torch._foreach_add([daily_infected, daily_deaths], [st.daily_infected, st.daily_deaths])
        return new_di, new_dd
    return ((daily_infected.reshape(-1).float() + st.daily_infected).to(daily_infected.dtype).view_as(daily_infected),
            (daily_deaths.reshape(-1).float() + st.daily_deaths).to(daily_deaths.dtype).view_as(daily_deaths))


@seirm_episode.register_fake
def _(handle, stage, next_stage_time, infected_time, age, daily_infected, daily_deaths, r_steps, m_value, num_steps, seed, iso_mode,
      p_isolate, isolation, uniform, want_grad, cuda_graph, want_table_grad):
    return torch.empty_like(daily_infected), torch.empty_like(daily_deaths)


@torch.library.custom_op("agenttorch_b200::seirm_episode_backward", mutates_args=())
def seirm_episode_backward(handle: int, num_steps: int, g_daily_infected: Tensor, g_daily_deaths: Tensor,
                           m_value: Tensor) -> Tensor:
    """(num_steps + 1,): dL/d(per-step R) followed by dL/dM (one buffer: one launch, one all-reduce on several GPUs)."""
    eng = get_engine(handle)
    sv = eng.saved_buffers(num_steps, True)
    return eng.episode_backward_packed(sv, num_steps, g_daily_infected, g_daily_deaths, m_value)


@seirm_episode_backward.register_fake
def _(handle, num_steps, g_daily_infected, g_daily_deaths, m_value):
    return g_daily_infected.new_empty(num_steps + 1)


class _SeirmEpisodeFn(torch.autograd.Function):
    """Differentiable episode: d(daily_infected, daily_deaths) / d(r_steps, M) by the fused backward kernel and, when the
    SFSusceptibility / SFInfector tables are given as differentiable tensors, d / d(tables) from the same saved buffers
    (SeirmEngine.table_gradients)."""

    @staticmethod
    def forward(ctx, handle, stage, next_stage_time, infected_time, age, daily_infected, daily_deaths, r_steps, m_value, sfs, sfi,
                num_steps, seed, iso_mode, p_isolate, isolation, uniform, want_grad, cuda_graph):
        table_grad = bool(want_grad and ((sfs is not None and sfs.requires_grad) or (sfi is not None and sfi.requires_grad)))
        di, dd = seirm_episode(handle, stage, next_stage_time, infected_time, age, daily_infected, daily_deaths,
                                     r_steps.detach(), m_value.detach(), num_steps, seed, iso_mode, p_isolate, isolation, uniform,
                                     want_grad, cuda_graph, table_grad)
        ctx.handle, ctx.num_steps, ctx.want_grad, ctx.table_grad = handle, num_steps, want_grad, table_grad
        ctx.r_shape, ctx.m_shape = r_steps.shape, m_value.shape
        ctx.sfs_shape = None if sfs is None else sfs.shape
        ctx.sfi_shape = None if sfi is None else sfi.shape
        ctx.save_for_backward(m_value.detach(), r_steps.detach(), age)
        return di, dd

    @staticmethod
    def backward(ctx, g_di, g_dd):
        m_value, r_steps, age = ctx.saved_tensors
        g_r = g_m = g_sfs = g_sfi = None
        if ctx.want_grad and (g_di is not None or g_dd is not None):
            T = ctx.num_steps
            zeros = torch.zeros(T, dtype=torch.float32, device=m_value.device) if (g_di is None or g_dd is None) else None
            gi = zeros if g_di is None else (g_di if g_di.dim() == 1 and g_di.numel() == T else g_di.reshape(-1)[:T])
            gd = zeros if g_dd is None else (g_dd if g_dd.dim() == 1 and g_dd.numel() == T else g_dd.reshape(-1)[:T])
            packed = seirm_episode_backward(ctx.handle, T, gi, gd, m_value)
            gr, gm = packed[:T], packed[T:]
            if r_steps.numel() == T:
                g_r = gr.reshape(ctx.r_shape)
            else:
                g_r = torch.zeros(ctx.r_shape, dtype=gr.dtype, device=gr.device).reshape(-1)
                g_r[:T] = gr
                g_r = g_r.reshape(ctx.r_shape)
            g_m = gm.reshape(ctx.m_shape) if m_value.numel() == 1 else gm.expand(m_value.numel()).reshape(ctx.m_shape).clone()
            if ctx.table_grad:
                eng = get_engine(ctx.handle)
                gs, gf = eng.table_gradients(eng.saved_buffers(T, True), T, gi, gd, m_value, r_steps, age)
                if ctx.sfs_shape is not None:
                    g_sfs = gs.reshape(ctx.sfs_shape)
                if ctx.sfi_shape is not None:
                    g_sfi = gf.reshape(ctx.sfi_shape)
        return (None, None, None, None, None, g_di, g_dd, g_r, g_m, g_sfs, g_sfi, None, None, None, None, None, None, None, None)


def seirm_episode_fn(handle, stage, next_stage_time, infected_time, age, daily_infected, daily_deaths, r_steps, m_value,
                     num_steps, seed, iso_mode="soft", p_isolate=0.0, isolation=None, uniform=None, want_grad=False, cuda_graph=False,
                     sfs=None, sfi=None):
    """The differentiable fused episode (see _SeirmEpisodeFn).  Returns (daily_infected, daily_deaths)."""
    return _SeirmEpisodeFn.apply(handle, stage, next_stage_time, infected_time, age, daily_infected, daily_deaths, r_steps, m_value,
                                 sfs, sfi, int(num_steps), int(seed), iso_mode, float(p_isolate), isolation, uniform, bool(want_grad),
                                 bool(cuda_graph))
