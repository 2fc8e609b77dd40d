"""Config helpers that wire the quarantine / testing substeps and the SEIRMS progression into a covid config.

The reference ships these classes (models/covid/substeps/{quarantine,testing,seirms_progression}/) and the state they
need (config.yaml:79-203,326-365) but neither registers them (models/covid/__init__.py:11-39) nor lists them under
`substeps:`.  This package registers them under the names below; `add_intervention_substeps(cfg)` appends the two
substep blocks (same block layout the reference's Initializer parses, core/initializer.py:319-388) and
`use_seirms_progression(cfg)` swaps the generator of substep '1'."""
import copy

CITIZEN = "agents/citizens/"
ENV = "environment/"


def _fn(generator, inputs, outputs, arguments=None):
    return {"generator": generator, "input_variables": inputs, "output_variables": outputs, "arguments": arguments}


def quarantine_substep():
    obs = {k: ENV + k for k in ("quarantine_start_prob", "quarantine_break_prob")}
    obs.update({k: CITIZEN + k for k in ("is_quarantined", "disease_stage")})
    return {
        "name": "Quarantine", "description": "Agents start, break or finish a quarantine", "active_agents": ["citizens"],
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
        "name": "Testing", "description": "Agents accept tests and receive results", "active_agents": ["citizens"],
        "observation": {"citizens": None},
        "policy": {"citizens": {"accept_test": _fn("AcceptTest", pol_in, ["test_acceptance_action"])}},
        "reward": None,
        "transition": {"update_test_status": _fn(
            "UpdateTestStatus", tr_in, ["awaiting_test_result", "test_result_date", "test_re_eligble_date"])},
    }


def add_intervention_substeps(config, quarantine=True, testing=True):
    """Returns a copy of `config` with the quarantine and/or testing substeps appended after the existing ones."""
    cfg = copy.deepcopy(config)
    subs = cfg["substeps"]
    for wanted, block in ((quarantine, quarantine_substep), (testing, testing_substep)):
        if wanted:
            subs[str(len(subs))] = block()
    cfg["simulation_metadata"]["num_substeps_per_step"] = len(subs)
    return cfg


def use_seirms_progression(config, substep="1", name="seirm_progression"):
    """Returns a copy of `config` whose progression transition is served by SEIRMSProgression: the block keeps its name,
    inputs, outputs and learnable M; only the registry entry it resolves to changes (`seirms_progression`)."""
    cfg = copy.deepcopy(config)
    block = cfg["substeps"][substep]["transition"].pop(name)
    block["generator"] = "SEIRMSProgression"
    cfg["substeps"][substep]["transition"]["seirms_progression"] = block
    return cfg


# ------------------------------------------------------------------------------------------ fused execution
def plan_intervention_substeps(runner, extra_keys):
    """For a config whose substeps after '0' (transmission) and '1' (progression) are exactly this package's quarantine and/or
    testing blocks: a factory `make_hook(state) -> hook(t)`; the Runner's fused SEIRM episode calls hook(t) right after the
    launches of step t, and the hook issues ONE in-place launch per block (policies + transition, no action tensors, no state
    copies) on the same stream.  None if any block is something else (then the generic Controller path runs)."""
    import re

    import torch

    from agenttorch_b200 import interventions as iv_ops
    from agenttorch_b200.core.helpers import get_by_path, set_by_path
    from agenttorch_b200.models.covid.substeps.common import ScalarCache
    from agenttorch_b200.models.covid.substeps.quarantine.action import BreakCompliance, StartCompliance
    from agenttorch_b200.models.covid.substeps.quarantine.observation import GetFromState
    from agenttorch_b200.models.covid.substeps.quarantine.transition import UpdateQuarantineStatus
    from agenttorch_b200.models.covid.substeps.testing.action import AcceptTest
    from agenttorch_b200.models.covid.substeps.testing.transition import UpdateTestStatus

    init = runner.initializer
    subs = runner.config["substeps"]
    tf, pf, of = init.transition_function, init.policy_function, init.observation_function
    scalar = ScalarCache()

    def modules(table, key, agent):
        return dict(table[key][agent]) if agent in table[key] else {}

    def exactly(mods, wanted):
        return list(mods.keys()) == [n for n, _ in wanted] and all(type(mods[n]) is c for n, c in wanted)

    def in_place(state, path):
        """float32 contiguous tensor at `path`, converted once and written back so the kernels can update it in place"""
        cur = get_by_path(state, re.split("/", path))
        fixed = cur if cur.dtype == torch.float32 and cur.is_contiguous() else cur.to(torch.float32).contiguous()
        if fixed is not cur:
            set_by_path(state, re.split("/", path), fixed)
        return fixed

    blocks = []
    for key in extra_keys:
        agents = subs[key]["active_agents"]
        if list(agents) != ["citizens"]:
            return None
        obs, pol, trn = modules(of, key, "citizens"), modules(pf, key, "citizens"), dict(tf[key])
        if (exactly(obs, [("get_from_state", GetFromState)]) and exactly(trn, [("update_quarantine_status", UpdateQuarantineStatus)])
                and sorted(pol) == ["break_compliance", "start_compliance"] and type(pol["start_compliance"]) is StartCompliance
                and type(pol["break_compliance"]) is BreakCompliance):
            blocks.append(("quarantine", obs["get_from_state"], pol["start_compliance"], pol["break_compliance"],
                           trn["update_quarantine_status"]))
        elif (not obs and exactly(pol, [("accept_test", AcceptTest)]) and exactly(trn, [("update_test_status", UpdateTestStatus)])):
            blocks.append(("testing", pol["accept_test"], trn["update_test_status"]))
        else:
            return None

    def still_fusable():
        """Conditions a user can change between calls: injected draws are only taken by the per-substep ops, and the fused
        quarantine launch draws both decisions from ONE hash per agent, so the two policies must share their noise address."""
        for block in blocks:
            if block[0] == "quarantine":
                start, brk = block[2], block[3]
                if start.uniform is not None or brk.uniform is not None:
                    return False
                if (start.noise_seed, start.advance_seed, start.episode_index) != (brk.noise_seed, brk.advance_seed, brk.episode_index):
                    return False
            elif block[1].uniform is not None:
                return False
        return True

    if not still_fusable():
        return None

    hook_cache = {}

    def make_hook(state):
        """hook(t): the in-place launches of step t.  hook.prepare(): once per episode, BEFORE the episode is issued or its CUDA
        graph replayed — draws the episode's noise seeds (the policies' counters advance) and writes them to device words the
        launches read (`seed_ptr`), so that a captured episode can be replayed with fresh noise.  hook.capture_key: everything a
        captured launch sequence bakes in (tensor addresses, probabilities, durations)."""
        calls, prepares, key = [], [], []
        for block in blocks:
            if block[0] == "quarantine":
                _, obs_mod, start, brk, upd = block
                iv, tv = obs_mod.input_variables, upd.input_variables
                stage = get_by_path(state, re.split("/", iv["disease_stage"]))
                is_q, date = in_place(state, tv["is_quarantined"]), in_place(state, tv["quarantine_start_date"])
                p_start = scalar("p_start", get_by_path(state, re.split("/", iv["quarantine_start_prob"])))
                p_break = scalar("p_break", get_by_path(state, re.split("/", iv["quarantine_break_prob"])))
                key += ["q", stage.data_ptr(), is_q.data_ptr(), date.data_ptr(), p_start, p_break, start.id_offset,
                        float(upd.quarantine_days), float(upd.INFINITY_TIME)]
            else:
                _, acc, upd = block
                pv, tv = acc.input_variables, upd.input_variables
                is_q = in_place(state, pv["is_quarantined"])
                aw, rd, re_ = (in_place(state, tv[k]) for k in ("awaiting_test_result", "test_result_date", "test_re_eligble_date"))
                p = scalar("p_comply", get_by_path(state, re.split("/", pv["test_compliance_prob"])))
                key += ["t", is_q.data_ptr(), aw.data_ptr(), rd.data_ptr(), re_.data_ptr(), p, acc.id_offset,
                        float(upd.test_ineligible_days), float(upd.test_result_delay_days), float(upd.INFINITY_TIME)]
        key = tuple(key)
        cached = hook_cache.get(key)
        if cached is not None:
            return cached
        for block in blocks:
            if block[0] == "quarantine":
                _, obs_mod, start, brk, upd = block
                iv, tv = obs_mod.input_variables, upd.input_variables
                stage = get_by_path(state, re.split("/", iv["disease_stage"]))
                is_q, date = in_place(state, tv["is_quarantined"]), in_place(state, tv["quarantine_start_date"])
                p_start = scalar("p_start", get_by_path(state, re.split("/", iv["quarantine_start_prob"])))
                p_break = scalar("p_break", get_by_path(state, re.split("/", iv["quarantine_break_prob"])))
                seed_q = torch.zeros(1, dtype=torch.int64, device=is_q.device)

                def prepare_quarantine(start=start, brk=brk, seed_q=seed_q):
                    seed, seed_b = start.episode_seed(0), brk.episode_seed(0)
                    assert seed == seed_b
                    seed_q.fill_(seed)

                def run_quarantine(t, stage=stage, is_q=is_q, date=date, start=start, upd=upd, p_start=p_start, p_break=p_break,
                                   seed_q=seed_q):
                    iv_ops.quarantine_step(stage.detach(), is_q, date, p_start, p_break, t, 0, start.id_offset,
                                           float(upd.quarantine_days), float(upd.INFINITY_TIME), seed_ptr=seed_q)
                prepares.append(prepare_quarantine)
                calls.append(run_quarantine)
            else:
                _, acc, upd = block
                pv, tv = acc.input_variables, upd.input_variables
                is_q = in_place(state, pv["is_quarantined"])
                aw, rd, re_ = (in_place(state, tv[k]) for k in ("awaiting_test_result", "test_result_date", "test_re_eligble_date"))
                p = scalar("p_comply", get_by_path(state, re.split("/", pv["test_compliance_prob"])))
                seed_t = torch.zeros(1, dtype=torch.int64, device=aw.device)

                def prepare_testing(acc=acc, seed_t=seed_t):
                    seed_t.fill_(acc.episode_seed(0))

                def run_testing(t, is_q=is_q, aw=aw, rd=rd, re_=re_, acc=acc, upd=upd, p=p, seed_t=seed_t):
                    iv_ops.testing_step(is_q, aw, rd, re_, p, t, 0, acc.id_offset, float(upd.test_ineligible_days),
                                        float(upd.test_result_delay_days), float(upd.INFINITY_TIME), seed_ptr=seed_t)
                prepares.append(prepare_testing)
                calls.append(run_testing)

        def hook(t):
            for call in calls:
                call(int(t))

        def prepare():
            for fn in prepares:
                fn()
        hook.launches_per_step = len(calls)
        hook.prepare = prepare
        hook.capture_key = key
        if len(hook_cache) > 4:
            hook_cache.clear()
        hook_cache[key] = hook
        return hook

    make_hook.still_fusable = still_fusable
    return make_hook
