"""Run the UNMODIFIED reference covid substeps that its registry does not wire up — quarantine, testing and
SEIRMS progression (models/covid/substeps/{quarantine,testing,seirms_progression}/) — on CPU, driven by the
reference's own Runner / Controller / Initializer.

TEST INFRASTRUCTURE (build container only, /root/reference mounted).  The classes are registered in the
reference's Registry under the names the extended YAML uses, the YAML copy gets two extra substeps (or the
SEIRMS class in place of `seirm_progression`), and the random draws are injected:
  * `discrete_sample` (core/helpers/soft.py:70-76, Gumbel-softmax tau=0.1 hard: index 0 with probability p)
    is replaced IN THE NAMESPACE of quarantine/action.py, testing/action.py and testing/transition.py by
    `(u < p)` on the counter-hash uniforms of oracle/noise.py (equal in law);
  * `SEIRMSProgression.st_bernoulli` (instance attribute, seirms_progression/transition.py:41) by `(u < p)`
    with identity backward, like `NewTransmission.st_bernoulli` in run_reference.py.
No reference source is copied into this repository.
"""
import os
import shutil
import tempfile

import numpy as np
import torch
import yaml

from oracle import noise
from oracle.ref_harness.run_reference import _InjectedBernoulli, _prepare_import, write_population

CITIZEN = "agents/citizens/"
ENV = "environment/"


def _fn(generator, inputs, outputs, arguments=None):
    return {"generator": generator, "input_variables": inputs, "output_variables": outputs, "arguments": arguments}


def quarantine_substep():
    obs = {k: ENV + k for k in ("quarantine_start_prob", "quarantine_break_prob")}
    obs.update({k: CITIZEN + k for k in ("is_quarantined", "disease_stage")})
    return {
        "name": "Quarantine", "description": "start / break / end quarantine", "active_agents": ["citizens"],
        "observation": {"citizens": {"get_from_state": _fn("GetFromState", obs, [])}},
        "policy": {"citizens": {
            "start_compliance": _fn("StartCompliance", {}, ["start_compliance_action"]),
            "break_compliance": _fn("BreakCompliance", {}, ["break_compliance_action"])}},
        "reward": None,
        "transition": {"update_quarantine_status": _fn(
            "UpdateQuarantineStatus",
            {"is_quarantined": CITIZEN + "is_quarantined", "quarantine_start_date": CITIZEN + "quarantine_start_date"},
            ["is_quarantined", "quarantine_start_date"])},
    }


def testing_substep():
    pol_in = {"is_quarantined": CITIZEN + "is_quarantined", "disease_stage": CITIZEN + "disease_stage",
              "test_compliance_prob": ENV + "test_compliance_prob"}
    tr_in = {"disease_stage": CITIZEN + "disease_stage", "test_result_date": CITIZEN + "test_result_date",
             "awaiting_test_result": CITIZEN + "awaiting_test_result", "test_re_eligble_date": CITIZEN + "test_re_eligble_date",
             "test_true_positive_prob": ENV + "test_true_positive_prob", "test_false_positive_prob": ENV + "test_false_positive_prob"}
    return {
        "name": "Testing", "description": "accept tests, receive results", "active_agents": ["citizens"],
        "observation": {"citizens": None},
        "policy": {"citizens": {"accept_test": _fn("AcceptTest", pol_in, ["test_acceptance_action"])}},
        "reward": None,
        "transition": {"update_test_status": _fn(
            "UpdateTestStatus", tr_in, ["awaiting_test_result", "test_result_date", "test_re_eligble_date"])},
    }


def run_interventions_reference(age, stage0, edges, seed, mode="hard", p_isolate=0.3, num_steps=21, variant="interventions",
                                meta=None, initial=None, env=None, r2=None, m=None, gI=None, gD=None):
    """variant 'interventions': substeps transmission, seirm_progression, quarantine, testing.
    variant 'seirms': transmission + SEIRMSProgression (registered as `seirm_progression`).
    meta: overrides of simulation_metadata (quarantine_days, test_*_days, RECOVERED_TO_SUSCEPTIBLE_TIME ...);
    initial: {agent property: scalar} initial values; env: {environment variable: scalar}."""
    work = _prepare_import()
    from agent_torch.core.dataloader import LoadPopulation
    from agent_torch.core.executor import Executor
    from agent_torch.models import covid
    from agent_torch.models.covid.substeps.quarantine import action as q_action
    from agent_torch.models.covid.substeps.quarantine import observation as q_obs
    from agent_torch.models.covid.substeps.quarantine import transition as q_transition
    from agent_torch.models.covid.substeps.seirm_progression.transition import SEIRMProgression
    from agent_torch.models.covid.substeps.seirms_progression.transition import SEIRMSProgression
    from agent_torch.models.covid.substeps.testing import action as t_action
    from agent_torch.models.covid.substeps.testing import transition as t_transition

    n = len(age)
    ids = np.arange(n)
    pop = tempfile.mkdtemp(prefix="pop_", dir=work)
    write_population(pop, age, stage0, edges)
    cfg_path = os.path.join(work, "agent_torch", "models", "covid", "yamls", "config.yaml")
    backup = cfg_path + ".orig"
    if not os.path.exists(backup):
        shutil.copy(cfg_path, backup)
    with open(backup) as f:
        cfg = yaml.safe_load(f)
    md = cfg["simulation_metadata"]
    md.update({"EXECUTION_MODE": "injected", "device": "cpu", "num_steps_per_episode": int(num_steps)})
    md.update(meta or {})
    for k, v in (initial or {}).items():
        cfg["state"]["agents"]["citizens"]["properties"][k]["value"] = v
    for k, v in (env or {}).items():
        cfg["state"]["environment"][k]["value"] = v
    reg = covid.registry
    if variant == "interventions":
        cfg["substeps"]["2"] = quarantine_substep()
        cfg["substeps"]["3"] = testing_substep()
        md["num_substeps_per_step"] = 4
        reg.register(SEIRMProgression, "seirm_progression", key="transition")
        reg.register(q_obs.GetFromState, "get_from_state", key="observation")
        reg.register(q_action.StartCompliance, "start_compliance", key="policy")
        reg.register(q_action.BreakCompliance, "break_compliance", key="policy")
        reg.register(q_transition.UpdateQuarantineStatus, "update_quarantine_status", key="transition")
        reg.register(t_action.AcceptTest, "accept_test", key="policy")
        reg.register(t_transition.UpdateTestStatus, "update_test_status", key="transition")
    elif variant == "seirms":
        reg.register(SEIRMSProgression, "seirm_progression", key="transition")
    else:
        raise ValueError(variant)
    with open(cfg_path, "w") as f:
        yaml.dump(cfg, f)

    runner = Executor(model=covid, pop_loader=LoadPopulation(pop)).runner
    runner.use_gpu = False
    runner.init()
    tr = runner.initializer.transition_function["0"]["new_transmission"]
    pg = runner.initializer.transition_function["1"]["seirm_progression"]
    pol = runner.initializer.policy_function["0"]["citizens"]["make_isolation_decision"]
    if r2 is not None:
        tr.calibrate_R2 = torch.tensor(np.asarray(r2, dtype=np.float32), requires_grad=True)
    if m is not None:
        if variant == "seirms":
            pg.external_M = torch.tensor(np.asarray(m, dtype=np.float32), requires_grad=True)
        else:
            pg.calibrate_M = torch.tensor(np.asarray(m, dtype=np.float32), requires_grad=True)

    def step_now():
        return int(runner.state["current_step"])

    pol.behavior = lambda obs: torch.from_numpy(noise.isolation_decision(seed, step_now(), ids, mode, p_isolate)).reshape(n, 1)
    tr.st_bernoulli = lambda probs: _InjectedBernoulli.apply(probs, torch.stack(
        (torch.from_numpy(noise.exposure_uniform(seed, step_now(), ids)).to(torch.float64), torch.zeros(n, dtype=torch.float64)), dim=1))

    calls = {"q": [None, 0], "t": [None, 0]}

    def injected(stream, key, per_step, order=(0, 1)):
        def sample(sample_prob, size, device="cpu", hard=True):
            t = step_now()
            slot = calls[key]
            if slot[0] != t:
                slot[0], slot[1] = t, 0
            k = slot[1]
            slot[1] += 1
            assert k < per_step
            u = noise.stream_uniforms(seed, t, ids, stream)[order[k]]
            p = (sample_prob * torch.ones(size)).reshape(-1)
            return (torch.from_numpy(u) < p).to(torch.float32)
        return sample

    saved = (q_action.discrete_sample, t_action.discrete_sample, t_transition.discrete_sample)
    mute = lambda *a, **k: None
    hist = {k: [] for k in ("disease_stage", "next_stage_time", "is_quarantined", "quarantine_start_date", "awaiting_test_result",
                             "test_result_date", "test_re_eligble_date")}
    try:
        if variant == "interventions":
            # the reference's DataLoader re-dumps the YAML with sorted keys (core/dataloader.py:58-62), so the policies run
            # alphabetically: call 0 = break_compliance (second uniform), call 1 = start_compliance (first uniform)
            q_action.discrete_sample = injected(noise.STREAM_QUARANTINE, "q", 2, order=(1, 0))
            t_action.discrete_sample = injected(noise.STREAM_TESTING, "t", 1)
            # the two draws of UpdateTestStatus only feed `positive_results`, which forward() drops (transition.py:119-160)
            t_transition.discrete_sample = lambda sample_prob, size, device="cpu", hard=True: torch.zeros(size)
            q_transition.print = mute
            t_action.print = mute
            t_transition.print = mute
        else:
            pg.st_bernoulli = lambda probs: _InjectedBernoulli.apply(probs, torch.from_numpy(
                noise.stream_uniforms(seed, step_now(), ids, noise.STREAM_MORTALITY)[0]).to(probs.dtype).reshape(probs.shape))
        runner.step(num_steps)
    finally:
        q_action.discrete_sample, t_action.discrete_sample, t_transition.discrete_sample = saved
        for mod in (q_transition, t_action, t_transition):
            mod.__dict__.pop("print", None)
        reg.register(SEIRMProgression, "seirm_progression", key="transition")
        shutil.copy(backup, cfg_path)
    for step_states in runner.state_trajectory[1:]:
        cit = step_states[-1]["agents"]["citizens"]
        for k in hist:
            hist[k].append(cit[k].detach().numpy().reshape(-1).astype(np.float64).copy())
    envs = runner.state_trajectory[-1][-1]["environment"]
    di, dd = envs["daily_infected"], envs["daily_deaths"]
    out = {k + "_history": np.stack(v) for k, v in hist.items()}
    out["daily_infected"] = di.detach().numpy().reshape(-1).copy()
    out["daily_deaths"] = dd.detach().numpy().reshape(-1).copy()
    out["dtypes"] = {k: str(runner.state["agents"]["citizens"][k].dtype) for k in hist}
    if gI is not None:
        loss = (torch.from_numpy(np.asarray(gI, dtype=np.float32)) * di.reshape(-1)).sum() + \
               (torch.from_numpy(np.asarray(gD, dtype=np.float32)) * dd.reshape(-1)).sum()
        loss.backward()
        out["loss"] = float(loss.detach())
        out["grad_R2"] = tr.calibrate_R2.grad.numpy().copy()
        mleaf = pg.external_M if variant == "seirms" else pg.calibrate_M
        out["grad_M"] = mleaf.grad.numpy().copy()
    shutil.rmtree(pop, ignore_errors=True)
    return out
