"""Emit agenttorch_b200/models/covid/yamls/config.yaml (same schema as the reference's covid config:
simulation_metadata / state / substeps; property = {name, shape, dtype, learnable, value | initialization_function})."""
import os

import yaml

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
N = "${simulation_metadata.num_agents}"
T = "${simulation_metadata.num_steps_per_episode}"


def prop(name, shape, dtype, value=None, gen=None, args=None, learnable=False):
    d = {"name": name, "shape": shape, "dtype": dtype, "learnable": learnable}
    if gen is None:
        d["initialization_function"] = None
        d["value"] = value
    else:
        d["initialization_function"] = {"generator": gen, "arguments": args or {}}
    return d


def arg(name, value, shape=(1,), learnable=False, dtype=None):
    d = {"name": name, "shape": list(shape) if shape is not None else None, "learnable": learnable,
         "initialization_function": None, "value": value}
    if dtype:
        d["dtype"] = dtype
    return d


stage_file = {"file_path": arg("CSV with one initial disease stage per agent", "${simulation_metadata.disease_stage_file}")}
age_file = {"file_path": arg("pickle with one age-group index per agent", "${simulation_metadata.age_group_file}")}
cfg = {
    "simulation_metadata": {
        "EXECUTION_MODE": "heuristic", "EXPOSED_TO_INFECTED_TIME": 5, "EXPOSED_VAR": 1, "INFECTED_TO_RECOVERED_TIME": 5,
        "INFECTED_VAR": 2, "INFINITY_TIME": 130, "MORTALITY_VAR": 4, "NUM_WEEKS": 3, "RECOVERED_VAR": 3,
        "RECOVERED_TO_SUSCEPTIBLE_TIME": 100, "quarantine_days": 12, "test_ineligible_days": 2, "test_result_delay_days": 3,
        "SUSCEPTIBLE_VAR": 0, "calibration": True, "device": "auto",
        "cuda_params": {"fused_episode": True, "cuda_graph": True, "snapshot_pack_bools": True, "batch_size": 16384,
                        "pool_limit_per_shape": 8, "inplace_progress": False},
        "population_dir": "populations/astoria",
        "age_group_file": "${simulation_metadata.population_dir}/age.pickle",
        "disease_stage_file": "${simulation_metadata.population_dir}/disease_stages.csv",
        "infection_network_file": "${simulation_metadata.population_dir}/mobility_networks/0.csv",
        "num_agents": 37518, "num_episodes": 5, "num_steps_per_episode": 21, "num_substeps_per_step": 2,
        "learning_params": {"lr": 5e-3, "betas": [0.5, 0.5], "lr_gamma": 0.9999, "model_path": "saved_model.pth"},
    },
    "state": {
        "agents": {"citizens": {"number": N, "properties": {
            "age": prop("age-group index 0..5", [N, 1], "int", gen="load_population_attribute",
                        args={"attribute": arg("attribute to load", "age", shape=None), **age_file}),
            "disease_stage": prop("S=0 E=1 I=2 R=3 (M=4)", [N, 1], "int", gen="read_from_file", args=stage_file),
            "id": prop("agent number", [N, 1], "float", gen="initialize_id",
                       args={"pop_size": arg("agents in the population", "${simulation_metadata.num_agents}")}),
            "infected_time": prop("step of the last exposure (500 = never)", [N, 1], "int", gen="get_infected_time",
                                  args={**stage_file, "num_steps": arg("steps in an episode", T)}),
            "next_stage_time": prop("step of the next stage change (500 = none pending)", [N, 1], "int", gen="get_next_stage_time",
                                    args={**stage_file, "num_steps": arg("steps in an episode", T)}),
            # state of the quarantine / testing substeps (reference config.yaml:79-87,145-203; float32 after `value * ones`)
            "awaiting_test_result": prop("1 while a test result is pending", [N, 1], "bool", value=False),
            "is_quarantined": prop("quarantine counter (1 = in quarantine)", [N, 1], "bool", value=False),
            "quarantine_start_date": prop("step at which the current quarantine began", [N, 1], "int",
                                          value="${simulation_metadata.INFINITY_TIME}"),
            "test_re_eligble_date": prop("first step at which the agent may test again", [N, 1], "int",
                                         value="${simulation_metadata.INFINITY_TIME}"),
            "test_result_date": prop("step at which the pending result arrives", [N, 1], "int",
                                     value="${simulation_metadata.INFINITY_TIME}"),
        }}},
        "environment": {
            "SFInfector": prop("infectiousness factor by disease stage of the source", [5], "float", value=[0.0, 0.33, 0.72, 0.0, 0.0]),
            "SFSusceptibility": prop("susceptibility factor by age group of the target", [9], "float",
                                     value=[0.35, 0.69, 1.03, 1.03, 1.03, 1.03, 1.27, 1.52, 1.52]),
            "daily_deaths": prop("deaths per step", [T], "float", value=0),
            "daily_infected": prop("new exposures per step", [T], "float", value=0),
            "exposed_to_infected_time": prop("E to I delay (unused by the substeps)", [1], "float", value=3),
            "infected_to_recovered_time": prop("I to R delay (unused by the substeps)", [1], "float", value=5),
            "lam_gamma_integrals": prop("infectiousness by days since exposure (gamma CDF differences)", [10], "float", gen="get_lam_gamma_integrals", args={
                "rate": arg("gamma rate", 2.4, dtype="float"), "scale": arg("gamma scale", 5.5, dtype="float"),
                "t": arg("number of days", T, dtype="int")}),
            "mean_interactions": prop("mean daily contacts by age group", [N, 1], "int",
                                      gen="get_mean_agent_interactions", args=age_file),
            "mortality_rate": prop("unused legacy mortality", [1], "float", value=0.2),
            "quarantine_break_prob": prop("P(break) per step while quarantined", [1], "float", value=0.1),
            "quarantine_start_prob": prop("P(start) per step while exposed or infected", [1], "float", value=0.7),
            "test_compliance_prob": prop("P(accept a test) per step", [1], "float", value=0.95),
            "test_false_positive_prob": prop("P(positive | healthy)", [1], "float", value=0.3),
            "test_true_positive_prob": prop("P(positive | exposed or infected)", [1], "float", value=0.8),
        },
        "network": {"agent_agent": {"infection_network": {
            "type": "network_from_file", "arguments": {"file_path": "${simulation_metadata.infection_network_file}"}}}},
        "objects": None,
    },
    "substeps": {
        "0": {
            "name": "Transmission", "description": "contact-graph exposure of susceptible agents", "active_agents": ["citizens"],
            "observation": {"citizens": None},
            "policy": {"citizens": {"make_isolation_decision": {
                "generator": "MakeIsolationDecision", "arguments": None,
                "input_variables": {"age": "agents/citizens/age"}, "output_variables": ["isolation_decision"]}}},
            "reward": None,
            "transition": {"new_transmission": {
                "generator": "NewTransmission",
                "arguments": {"R2": {"name": "R2", "dtype": "float", "shape": ["${simulation_metadata.NUM_WEEKS}", 1],
                                     "learnable": True, "initialization_function": None, "value": 4.75}},
                "input_variables": {
                    "SFInfector": "environment/SFInfector", "SFSusceptibility": "environment/SFSusceptibility",
                    "adjacency_matrix": "network/agent_agent/infection_network/adjacency_matrix",
                    "age": "agents/citizens/age", "daily_infected": "environment/daily_infected",
                    "disease_stage": "agents/citizens/disease_stage",
                    "exposed_to_infected_time": "environment/exposed_to_infected_time",
                    "infected_time": "agents/citizens/infected_time",
                    "lam_gamma_integrals": "environment/lam_gamma_integrals",
                    "mean_interactions": "environment/mean_interactions",
                    "next_stage_time": "agents/citizens/next_stage_time"},
                "output_variables": ["disease_stage", "next_stage_time", "infected_time", "daily_infected"]}},
        },
        "1": {
            "name": "Disease Progression", "description": "E to I to R on the stage timers, daily deaths",
            "active_agents": ["citizens"], "observation": {"citizens": None}, "policy": {"citizens": None}, "reward": None,
            "transition": {"seirm_progression": {
                "generator": "SEIRMProgression",
                "arguments": {"M": {"name": "mortality M", "dtype": "float", "shape": [1], "learnable": True,
                                    "initialization_function": None, "value": 0.12}},
                "input_variables": {
                    "daily_deaths": "environment/daily_deaths", "disease_stage": "agents/citizens/disease_stage",
                    "exposed_to_infected_time": "environment/exposed_to_infected_time",
                    "infected_to_recovered_time": "environment/infected_to_recovered_time",
                    "next_stage_time": "agents/citizens/next_stage_time"},
                "output_variables": ["disease_stage", "next_stage_time", "daily_deaths"]}},
        },
    },
}
out = os.path.join(ROOT, "agenttorch_b200", "models", "covid", "yamls", "config.yaml")
with open(out, "w") as f:
    f.write("# SEIRM covid model — same schema as the reference's models/covid/yamls/config.yaml; generated by tools/gen_covid_config.py\n")
    yaml.safe_dump(cfg, f, sort_keys=False, default_flow_style=None, width=118)
print("wrote", out)
