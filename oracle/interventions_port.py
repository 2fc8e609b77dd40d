"""CPU restatement ("port") of the covid substeps the reference ships but does not register — quarantine,
testing, SEIRMS progression — the parity oracle for csrc/at_interventions.cu.

TEST INFRASTRUCTURE ONLY (tests/, smoke(); never imported by agenttorch_b200/).

Pinned against the real reference: tests/golden/interventions_*.npz and seirms_*.npz are outputs of
/root/reference itself (oracle/ref_harness/make_golden_interventions.py, through the reference's own Runner);
tests/test_oracle_interventions.py checks this port against them bit for bit (per-step histories of every
property) and to 1e-6 for the series and gradients.  The reference's own tests never touch these classes.

Restated, per agent and in closed form (paths relative to /root/reference/agent_torch/models/covid/substeps):
  quarantine/action.py:16-68        StartCompliance / BreakCompliance        -> quarantine_step
  quarantine/transition.py:31-122   UpdateQuarantineStatus (end, start, break)
  testing/action.py:29-53           AcceptTest                               -> testing_step
  testing/transition.py:41-175      UpdateTestStatus (results, enrolment)
  seirms_progression/transition.py:48-181  SEIRMSProgression                 -> SeirmsPort.progression
Quirks kept on purpose (I1-I7, documented in DESIGN.md section 8):
  I1 `_end_quarantine` does not look at is_quarantined, and a break decision is taken from the status observed
     BEFORE the transition: an agent whose quarantine ends and breaks in the same step goes to -1, and a non-zero
     status keeps "breaking", so is_quarantined drifts below zero (it is a float32 counter, not a flag);
  I2 "not quarantined" is `is_quarantined.long() == 0` (soft.py:27-38);
  I3 AcceptTest's eligibility is `stage > S OR stage < R`, true for every stage (action.py:38-41), so only
     quarantine status and the compliance draw decide;
  I4 the true/false-positive draws of UpdateTestStatus feed a value its forward() drops: no observable effect;
  I5 SEIRMSProgression.forward unpacks (deaths, dead, recovered) as (deaths, recovered, dead): agents the
     Bernoulli draw killed move I -> R (stage 3), survivors move I -> stage 4;
  I6 `new_transition_times[newly_dead.long()] = ...` indexes ROWS 0 and 1 by the 0/1 VALUES: the timers of agents
     0 and 1 are overwritten every step, nobody else's recovery timer is ever set, so a stage-3 agent is "due" at
     once and bounces R -> S on the next step;
  I7 stage transitions are written `stage + (stage/k) * move`, so d new_stage / d stage is 2 across E -> I,
     0 across R -> S and 3 - dead/2 across I -> R/M: gradients w.r.t. R2 through daily_deaths double.
"""
import numpy as np
import torch

from oracle import noise
from oracle.seirm_port import SeirmPort, SeirmState


# ------------------------------------------------------------------------------------------ quarantine
def quarantine_step(t, stage, is_q, start_date, u_start, u_break, p_start, p_break, quarantine_days, infinity_time):
    """One quarantine substep.  All arrays (N,) float32 (stage any float); returns (is_q, start_date) float32."""
    stage = np.asarray(stage, dtype=np.float64)
    is_q = np.asarray(is_q, dtype=np.float32)
    sd = np.asarray(start_date, dtype=np.float32)
    not_q = np.trunc(is_q) == 0                                                         # I2
    start = (u_start < np.float32(p_start)) & not_q & ((stage == 1) | (stage == 2))     # action.py:27-46
    brk = (is_q != 0) & (u_break < np.float32(p_break))                                 # action.py:59-66
    ends = np.float32(t) >= sd + np.float32(quarantine_days)                            # transition.py:31-33 (I1)
    f = np.float32
    new_q = is_q + ends.astype(f) * f(-1) + start.astype(f) - brk.astype(f)
    sd = sd * (1 - ends.astype(f)) + f(infinity_time) * ends.astype(f)
    sd = sd * (~start).astype(f) + start.astype(f) * f(t)
    sd = sd * (~brk).astype(f) + brk.astype(f) * f(infinity_time)
    return new_q.astype(f), sd.astype(f)


# --------------------------------------------------------------------------------------------- testing
def testing_step(t, is_q, awaiting, result_date, re_eligible_date, u_comply, p_comply, ineligible_days, delay_days,
                 infinity_time):
    """One testing substep; returns (awaiting, result_date, re_eligible_date) float32."""
    f = np.float32
    is_q = np.asarray(is_q, dtype=f)
    aw = np.asarray(awaiting, dtype=f).copy()
    rd = np.asarray(result_date, dtype=f).copy()
    re = np.asarray(re_eligible_date, dtype=f).copy()
    willing = (np.trunc(is_q) == 0) & (u_comply < f(p_comply))                           # I3
    today = rd == f(t)                                                                   # transition.py:53-58
    aw = aw - today.astype(f)
    re[today] = f(t + ineligible_days)
    rd[today] = f(infinity_time)
    ineligible = (f(t) < re) | (np.trunc(aw) != 0)                                       # transition.py:152-156
    enrolled = willing & ~ineligible
    aw = aw + enrolled.astype(f)
    rd[enrolled] = f(t + delay_days)
    return aw, rd, re


# ---------------------------------------------------------------------------------------------- SEIRMS
class SeirmsPort(SeirmPort):
    """SeirmPort with SEIRMSProgression as substep '1' (torch autograd gives the reference's gradients)."""

    def __init__(self, *args, recovered_to_susceptible_time=100, **kwargs):
        super().__init__(*args, **kwargs)
        self.r2s = recovered_to_susceptible_time

    def progression_seirms(self, st, t, m, u_death):
        c = self.c
        stage, nst = st.stage, st.next_stage_time
        due = nst <= t
        rod = ((stage == c.infected) & due)
        ndr = stage * rod / c.infected
        deaths = ndr.sum() * m
        daily_deaths = st.daily_deaths + torch.eye(c.num_steps)[t] * deaths
        prob = m * rod                                                     # float32 (N,)
        hit = (torch.as_tensor(u_death, dtype=torch.float32) < prob.detach()).to(prob.dtype)
        dead = prob + (hit - prob).detach()                                # straight-through
        recovered = ndr - dead
        to_inf = ((stage == c.exposed) & due) * 1
        to_sus = ((stage == c.recovered) & due) * (-c.recovered)
        # I5: the survivors take the "dead" path (two levels up), the dead the "recovered" path (one level up)
        move = (stage / c.exposed) * to_inf + (stage / c.infected) * (recovered * 2) + (stage / c.infected) * dead \
            + (stage / c.recovered) * to_sus
        new_stage = stage + move
        new_nst = nst.clone()
        new_nst[(stage.detach() == c.exposed) & (nst == t)] = float(t + c.infected_to_recovered_time)
        new_nst[recovered.detach().long()] = float(c.infinity_time)        # I6: rows 0 / 1
        new_nst[dead.detach().long()] = float(self.r2s)
        return SeirmState(new_stage, new_nst, st.infected_time, st.daily_infected, daily_deaths.reshape(-1))


def interventions_episode(port, seed, mode, p_isolate, num_steps, r2, m, meta, initial, env):
    """Transmission + SEIRM progression + quarantine + testing for num_steps; per-step histories (numpy)."""
    n = port.n
    ids = np.arange(n)
    st = port.initial_state()
    f = np.float32
    inf_t = meta["INFINITY_TIME"]
    is_q = np.full(n, initial.get("is_quarantined", 0), dtype=f)
    qsd = np.full(n, initial.get("quarantine_start_date", inf_t), dtype=f)
    aw = np.full(n, initial.get("awaiting_test_result", 0), dtype=f)
    rd = np.full(n, initial.get("test_result_date", inf_t), dtype=f)
    re = np.full(n, initial.get("test_re_eligble_date", inf_t), dtype=f)
    hist = {k: [] for k in ("disease_stage", "next_stage_time", "is_quarantined", "quarantine_start_date",
                             "awaiting_test_result", "test_result_date", "test_re_eligble_date")}
    r2 = torch.as_tensor(np.asarray(r2, dtype=f))
    m = torch.as_tensor(np.asarray(m, dtype=f))
    for t in range(num_steps):
        w = torch.from_numpy(noise.isolation_decision(seed, t, ids, mode, p_isolate))
        u = torch.from_numpy(noise.exposure_uniform(seed, t, ids))
        st, _ = port.transmission(st, t, w, u, r2)
        st = port.progression(st, t, m)
        us, ub = noise.stream_uniforms(seed, t, ids, noise.STREAM_QUARANTINE)
        is_q, qsd = quarantine_step(t, st.stage.numpy(), is_q, qsd, us, ub, env["quarantine_start_prob"],
                                    env["quarantine_break_prob"], meta["quarantine_days"], inf_t)
        uc, _ = noise.stream_uniforms(seed, t, ids, noise.STREAM_TESTING)
        aw, rd, re = testing_step(t, is_q, aw, rd, re, uc, env["test_compliance_prob"], meta["test_ineligible_days"],
                                  meta["test_result_delay_days"], inf_t)
        for k, v in zip(hist, (st.stage.numpy(), st.next_stage_time.numpy(), is_q, qsd, aw, rd, re)):
            hist[k].append(np.asarray(v, dtype=np.float64).copy())
    out = {k + "_history": np.stack(v) for k, v in hist.items()}
    out["daily_infected"] = st.daily_infected.numpy()
    out["daily_deaths"] = st.daily_deaths.numpy()
    return out


def seirms_episode_with_grads(port, seed, mode, p_isolate, num_steps, r2, m, gI, gD):
    n = port.n
    ids = np.arange(n)
    r2 = torch.tensor(np.asarray(r2, dtype=np.float32), requires_grad=True)
    m = torch.tensor(np.asarray(m, dtype=np.float32), requires_grad=True)
    st = port.initial_state()
    stages, timers = [], []
    for t in range(num_steps):
        w = torch.from_numpy(noise.isolation_decision(seed, t, ids, mode, p_isolate))
        u = torch.from_numpy(noise.exposure_uniform(seed, t, ids))
        st, _ = port.transmission(st, t, w, u, r2)
        st = port.progression_seirms(st, t, m, noise.stream_uniforms(seed, t, ids, noise.STREAM_MORTALITY)[0])
        stages.append(st.stage.detach().numpy().copy())
        timers.append(st.next_stage_time.numpy().astype(np.float64).copy())
    loss = (torch.as_tensor(np.asarray(gI, dtype=np.float32)) * st.daily_infected).sum() + \
           (torch.as_tensor(np.asarray(gD, dtype=np.float32)) * st.daily_deaths).sum()
    loss.backward()
    return {"disease_stage_history": np.stack(stages), "next_stage_time_history": np.stack(timers),
            "daily_infected": st.daily_infected.detach().numpy(), "daily_deaths": st.daily_deaths.detach().numpy(),
            "loss": float(loss.detach()), "grad_R2": r2.grad.numpy().copy(), "grad_M": m.grad.numpy().copy()}
