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
