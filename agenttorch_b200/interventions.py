"""torch.library custom ops (namespace `agenttorch_b200`) over libat_interventions.so: the covid model's quarantine,
testing and SEIRMS-progression substeps (reference classes in models/covid/substeps/{quarantine,testing,
seirms_progression}/, shipped but not registered there).

  quarantine_decide   StartCompliance / BreakCompliance.forward      (quarantine/action.py:26-68)
  quarantine_update   UpdateQuarantineStatus.forward                  (quarantine/transition.py:31-122)
  quarantine_step     both in one launch, no action tensors materialised
  testing_accept      AcceptTest.forward                              (testing/action.py:29-53)
  testing_update      UpdateTestStatus.forward                        (testing/transition.py:41-175)
  testing_step        both in one launch
  seirms_progression  SEIRMSProgression.forward, with registered autograd (seirms_progression/transition.py:48-181)
Every op launches CUDA kernels through the C ABI (include/at_interventions.h); there is no CPU implementation —
CPU tensors raise.
"""
from typing import Optional, Tuple

import torch
from torch import Tensor

from . import _native_iv as nat


def _ptr(t):
    return None if t is None else t.data_ptr()


def _stream(t):
    return torch.cuda.current_stream(t.device).cuda_stream


def _require_cuda(name, t):
    """No CPU fallback.  (One function so that the host-emulation tests, which drive these ops against the kernel source
    compiled for the CPU, can lift the check.)"""
    if not t.is_cuda:
        raise RuntimeError(f"agenttorch_b200 intervention kernels need CUDA tensors ({name} is on {t.device}); there is no CPU fallback")


def _f32(t, name):
    if t is None:
        return None
    _require_cuda(name, t)
    return t if t.dtype == torch.float32 and t.is_contiguous() else t.to(torch.float32).contiguous()


def _u64(seed):
    return int(seed) & ((1 << 64) - 1)


def _as_flag_bytes(t):
    """Action tensors written by other policies may be bool, integer or float: non-zero means 'yes'."""
    t = t.reshape(-1)
    if t.dtype == torch.bool:
        return t.contiguous().view(torch.uint8)
    if t.dtype == torch.uint8:
        return t.contiguous()
    return (t != 0).view(torch.uint8)


def _quarantine(flags, t, seed, id_offset, p_start, p_break, days, infinity, stage, is_q, start_date, u_start, u_break,
                start_action, break_action, is_q_out, start_date_out, seed_ptr=None):
    n = is_q.numel()
    stage_f64 = 0
    if stage is not None:
        if stage.dtype == torch.float64:
            stage_f64, stage = 1, stage.contiguous()
        else:
            stage = _f32(stage, "disease_stage")
    a = nat.QuarantineArgs(
        abi_version=nat.ABI_VERSION, flags=flags, t=int(t), stage_f64=stage_f64, n=n, id_offset=int(id_offset), seed=_u64(seed),
        seed_ptr=_ptr(seed_ptr), p_start=float(p_start), p_break=float(p_break), quarantine_days=float(days), infinity_time=float(infinity),
        stage=_ptr(stage), u_start=_ptr(u_start), u_break=_ptr(u_break), is_q_in=_ptr(is_q), start_date_in=_ptr(start_date),
        start_action=_ptr(start_action), break_action=_ptr(break_action), is_q_out=_ptr(is_q_out), start_date_out=_ptr(start_date_out))
    nat.check(nat.lib().at_iv_quarantine(a, _stream(is_q)))


@torch.library.custom_op("agenttorch_b200::quarantine_decide", mutates_args=())
def quarantine_decide(stage: Tensor, is_quarantined: Tensor, p_start: float, p_break: float, t: int, seed: int, id_offset: int,
                      u_start: Optional[Tensor], u_break: Optional[Tensor]) -> Tuple[Tensor, Tensor]:
    """(start_compliance_action, break_compliance_action): bool tensors shaped like is_quarantined."""
    is_q = _f32(is_quarantined, "is_quarantined").reshape(-1)
    start = torch.empty(is_q.numel(), dtype=torch.bool, device=is_q.device)
    brk = torch.empty_like(start)
    _quarantine(nat.DECIDE, t, seed, id_offset, p_start, p_break, 0.0, 0.0, stage.reshape(-1), is_q, None,
                _f32(u_start, "u_start"), _f32(u_break, "u_break"), start, brk, None, None)
    return start.view(is_quarantined.shape), brk.view(is_quarantined.shape)


@quarantine_decide.register_fake
def _(stage, is_quarantined, p_start, p_break, t, seed, id_offset, u_start, u_break):
    e = torch.empty_like(is_quarantined, dtype=torch.bool)
    return e, torch.empty_like(e)


@torch.library.custom_op("agenttorch_b200::quarantine_update", mutates_args=())
def quarantine_update(is_quarantined: Tensor, quarantine_start_date: Tensor, start_action: Tensor, break_action: Tensor,
                      t: int, quarantine_days: float, infinity_time: float) -> Tuple[Tensor, Tensor]:
    is_q = _f32(is_quarantined, "is_quarantined").reshape(-1)
    sd = _f32(quarantine_start_date, "quarantine_start_date").reshape(-1)
    out_q, out_d = torch.empty_like(is_q), torch.empty_like(sd)
    _quarantine(nat.APPLY, t, 0, 0, 0.0, 0.0, quarantine_days, infinity_time, None, is_q, sd, None, None,
                _as_flag_bytes(start_action), _as_flag_bytes(break_action), out_q, out_d)
    return out_q.view(is_quarantined.shape), out_d.view(quarantine_start_date.shape)


@quarantine_update.register_fake
def _(is_quarantined, quarantine_start_date, start_action, break_action, t, quarantine_days, infinity_time):
    return (torch.empty_like(is_quarantined, dtype=torch.float32), torch.empty_like(quarantine_start_date, dtype=torch.float32))


@torch.library.custom_op("agenttorch_b200::quarantine_step", mutates_args=("is_quarantined", "quarantine_start_date"))
def quarantine_step(stage: Tensor, is_quarantined: Tensor, quarantine_start_date: Tensor, p_start: float, p_break: float,
                    t: int, seed: int, id_offset: int, quarantine_days: float, infinity_time: float, seed_ptr: Optional[Tensor] = None) -> None:
    """Policies + transition of one quarantine substep IN PLACE (float32 contiguous state), one launch, untouched 16-byte
    groups are not written.  seed_ptr: one-element int64 device tensor that overrides `seed` (a captured CUDA graph is replayed
    with another seed by rewriting it)."""
    for x in (is_quarantined, quarantine_start_date):
        _require_cuda("in-place state", x)
        if x.dtype != torch.float32 or not x.is_contiguous():
            raise RuntimeError("quarantine_step works in place on contiguous float32 CUDA tensors")
    _quarantine(nat.DECIDE | nat.APPLY, t, seed, id_offset, p_start, p_break, quarantine_days, infinity_time, stage.reshape(-1),
                is_quarantined.view(-1), quarantine_start_date.view(-1), None, None, None, None, is_quarantined.view(-1),
                quarantine_start_date.view(-1), seed_ptr=seed_ptr)


def _testing(flags, t, seed, id_offset, p_comply, inelig, delay, infinity, u_comply, is_q, accept, aw, rd, re, aw_o, rd_o, re_o,
             seed_ptr=None):
    ref = is_q if is_q is not None else aw
    a = nat.TestingArgs(
        abi_version=nat.ABI_VERSION, flags=flags, t=int(t), reserved=0, n=ref.numel(), id_offset=int(id_offset), seed=_u64(seed),
        seed_ptr=_ptr(seed_ptr), p_comply=float(p_comply), ineligible_days=float(inelig), result_delay_days=float(delay),
        infinity_time=float(infinity), u_comply=_ptr(u_comply), is_q=_ptr(is_q), accept_action=_ptr(accept),
        awaiting_in=_ptr(aw), result_date_in=_ptr(rd), re_eligible_in=_ptr(re), awaiting_out=_ptr(aw_o), result_date_out=_ptr(rd_o),
        re_eligible_out=_ptr(re_o))
    nat.check(nat.lib().at_iv_testing(a, _stream(ref)))


@torch.library.custom_op("agenttorch_b200::testing_accept", mutates_args=())
def testing_accept(is_quarantined: Tensor, p_comply: float, t: int, seed: int, id_offset: int, u_comply: Optional[Tensor]) -> Tensor:
    """test_acceptance_action: float32 0/1 shaped like is_quarantined."""
    is_q = _f32(is_quarantined, "is_quarantined").reshape(-1)
    out = torch.empty_like(is_q)
    _testing(nat.DECIDE, t, seed, id_offset, p_comply, 0.0, 0.0, 0.0, _f32(u_comply, "u_comply"), is_q, out,
             None, None, None, None, None, None)
    return out.view(is_quarantined.shape)


@testing_accept.register_fake
def _(is_quarantined, p_comply, t, seed, id_offset, u_comply):
    return torch.empty_like(is_quarantined, dtype=torch.float32)


@torch.library.custom_op("agenttorch_b200::testing_update", mutates_args=())
def testing_update(awaiting_test_result: Tensor, test_result_date: Tensor, test_re_eligble_date: Tensor, accept_action: Tensor,
                   t: int, ineligible_days: float, result_delay_days: float, infinity_time: float) -> Tuple[Tensor, Tensor, Tensor]:
    aw = _f32(awaiting_test_result, "awaiting_test_result").reshape(-1)
    rd = _f32(test_result_date, "test_result_date").reshape(-1)
    re = _f32(test_re_eligble_date, "test_re_eligble_date").reshape(-1)
    acc = _f32(accept_action, "test_acceptance_action").reshape(-1)
    o = [torch.empty_like(x) for x in (aw, rd, re)]
    _testing(nat.APPLY, t, 0, 0, 0.0, ineligible_days, result_delay_days, infinity_time, None, None, acc, aw, rd, re, *o)
    return (o[0].view(awaiting_test_result.shape), o[1].view(test_result_date.shape), o[2].view(test_re_eligble_date.shape))


@testing_update.register_fake
def _(awaiting_test_result, test_result_date, test_re_eligble_date, accept_action, t, ineligible_days, result_delay_days, infinity_time):
    return tuple(torch.empty_like(x, dtype=torch.float32) for x in (awaiting_test_result, test_result_date, test_re_eligble_date))


@torch.library.custom_op("agenttorch_b200::testing_step",
                         mutates_args=("awaiting_test_result", "test_result_date", "test_re_eligble_date"))
def testing_step(is_quarantined: Tensor, awaiting_test_result: Tensor, test_result_date: Tensor, test_re_eligble_date: Tensor,
                 p_comply: float, t: int, seed: int, id_offset: int, ineligible_days: float, result_delay_days: float,
                 infinity_time: float, seed_ptr: Optional[Tensor] = None) -> None:
    """Policy + transition of one testing substep IN PLACE, one launch (seed_ptr: see quarantine_step)."""
    for x in (awaiting_test_result, test_result_date, test_re_eligble_date):
        _require_cuda("in-place state", x)
        if x.dtype != torch.float32 or not x.is_contiguous():
            raise RuntimeError("testing_step works in place on contiguous float32 CUDA tensors")
    aw, rd, re = awaiting_test_result.view(-1), test_result_date.view(-1), test_re_eligble_date.view(-1)
    _testing(nat.DECIDE | nat.APPLY, t, seed, id_offset, p_comply, ineligible_days, result_delay_days, infinity_time, None,
             _f32(is_quarantined, "is_quarantined").reshape(-1), None, aw, rd, re, aw, rd, re, seed_ptr=seed_ptr)


# ------------------------------------------------------------------------------------------ SEIRMS progression
# d new_stage / d stage per move code (oracle/interventions_port.py quirk I7): none 1, E->I 2, R->S 0, survived 3, died 2.5
_STAGE_FACTOR = (1.0, 2.0, 0.0, 3.0, 2.5)


@torch.library.custom_op("agenttorch_b200::seirms_progression", mutates_args=())
def seirms_progression(stage: Tensor, next_stage_time: Tensor, daily_deaths: Tensor, m_value: Tensor, t: int, seed: int,
                       id_offset: int, u_death: Optional[Tensor], infected_to_recovered_time: float, infinity_time: float,
                       recovered_to_susceptible_time: float) -> Tuple[Tensor, Tensor, Tensor, Tensor, Tensor]:
    """Returns (stage, next_stage_time, daily_deaths, move_code uint8 (N,), counts int64 [due infected, died, survived, 0])."""
    s = _f32(stage.detach(), "disease_stage").reshape(-1)
    nt = _f32(next_stage_time.detach(), "next_stage_time").reshape(-1)
    dd = _f32(daily_deaths.detach(), "daily_deaths").reshape(-1).clone()
    m = _f32(m_value.detach(), "M").reshape(-1)
    if t >= dd.numel():
        raise IndexError(f"step {t} outside daily_deaths of length {dd.numel()}")
    s_out, nt_out = torch.empty_like(s), torch.empty_like(nt)
    code = torch.empty(s.numel(), dtype=torch.uint8, device=s.device)
    counts = torch.zeros(4, dtype=torch.int64, device=s.device)
    a = nat.SeirmsArgs(
        abi_version=nat.ABI_VERSION, t=int(t), n=s.numel(), id_offset=int(id_offset), seed=_u64(seed), seed_ptr=None,
        infected_to_recovered_time=float(infected_to_recovered_time), infinity_time=float(infinity_time),
        recovered_to_susceptible_time=float(recovered_to_susceptible_time), num_steps=dd.numel(), m_value=_ptr(m),
        u_death=_ptr(_f32(u_death, "u_death")), stage_in=_ptr(s), next_stage_time_in=_ptr(nt), stage_out=_ptr(s_out),
        next_stage_time_out=_ptr(nt_out), move_code=_ptr(code), counts=_ptr(counts), daily_deaths=_ptr(dd))
    nat.check(nat.lib().at_iv_seirms_progression(a, _stream(s)))
    return s_out.view(stage.shape), nt_out.view(next_stage_time.shape), dd.view(daily_deaths.shape), code, counts


@seirms_progression.register_fake
def _(stage, next_stage_time, daily_deaths, m_value, t, seed, id_offset, u_death, infected_to_recovered_time, infinity_time,
      recovered_to_susceptible_time):
    return (torch.empty_like(stage, dtype=torch.float32), torch.empty_like(next_stage_time, dtype=torch.float32),
            torch.empty_like(daily_deaths, dtype=torch.float32), stage.new_empty(stage.numel(), dtype=torch.uint8),
            stage.new_empty(4, dtype=torch.int64))


def seirms_backward_formula(code, counts, m_value, t, g_stage, g_dd, stage_shape):
    """Gradients of one SEIRMS progression call w.r.t. (stage, M) from the saved move codes (any device)."""
    g_stage_in = g_m = None
    due = (code >= nat.MOVE_SURVIVED).view(stage_shape)
    if g_stage is not None:
        lut = torch.tensor(_STAGE_FACTOR, dtype=g_stage.dtype, device=code.device)
        g_stage_in = g_stage * lut[code.long()].view(stage_shape)
        # straight-through death draw: d dead / dM = [due infected], d new_stage / d dead = -stage/2 = -1
        g_m = -(g_stage * due.to(g_stage.dtype)).sum()
    if g_dd is not None:
        gt = g_dd.reshape(-1)[t]
        # daily_deaths[t] += M * sum(stage * due / 2)
        add = (gt * m_value.reshape(-1)[0].to(gt.dtype) * 0.5) * due.to(gt.dtype)
        g_stage_in = add if g_stage_in is None else g_stage_in + add.to(g_stage_in.dtype)
        gm2 = gt * counts[0].to(gt.dtype)
        g_m = gm2 if g_m is None else g_m.to(gm2.dtype) + gm2
    return g_stage_in, g_m


def _seirms_setup(ctx, inputs, output):
    ctx.t = inputs[4]
    ctx.m_shape, ctx.stage_shape = inputs[3].shape, inputs[0].shape
    ctx.mark_non_differentiable(output[1], output[3], output[4])
    ctx.save_for_backward(output[3], output[4], inputs[3])


def _seirms_backward(ctx, g_stage, g_nst, g_dd, g_code, g_counts):
    code, counts, m = ctx.saved_tensors
    g_stage_in, g_m = seirms_backward_formula(code, counts, m, ctx.t, g_stage, g_dd, ctx.stage_shape)
    if g_m is not None:
        g_m = g_m.reshape(1).expand(m.numel()).reshape(ctx.m_shape).to(m.dtype).clone()
    return (g_stage_in, None, g_dd, g_m, None, None, None, None, None, None, None)


seirms_progression.register_autograd(_seirms_backward, setup_context=_seirms_setup)
