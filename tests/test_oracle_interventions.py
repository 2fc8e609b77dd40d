"""oracle/interventions_port.py against the outputs of the real reference (tests/golden/interventions_*.npz,
seirms_*.npz, written by oracle/ref_harness/make_golden_interventions.py): per-step histories bit for bit, daily
series and gradients to 1e-6."""
import os

import numpy as np
import pytest

from oracle.interventions_port import SeirmsPort, interventions_episode, seirms_episode_with_grads
from oracle.seirm_port import SeirmPort
from tests.helpers import GOLDEN_DIR, INTERVENTION_CASES, SEIRMS_CASES, load_interventions_golden, rel_err


@pytest.mark.parametrize("name", INTERVENTION_CASES)
def test_quarantine_and_testing_match_reference(name):
    g = load_interventions_golden(name)
    ov = g["overrides"]
    out = interventions_episode(SeirmPort(g["age"], g["stage0"], g["edges"]), g["seed"], g["mode"], float(g["p_isolate"]),
                                g["num_steps"], g["r2"], g["m"], ov["meta"], ov["initial"], ov["env"])
    for k in ("disease_stage", "next_stage_time", "is_quarantined", "quarantine_start_date", "awaiting_test_result",
              "test_result_date", "test_re_eligble_date"):
        assert np.array_equal(out[k + "_history"], g["ref_" + k + "_history"].astype(np.float64)), k
    assert rel_err(g["ref_daily_infected"], out["daily_infected"]) <= 1e-6
    assert rel_err(g["ref_daily_deaths"], out["daily_deaths"]) <= 1e-6


@pytest.mark.parametrize("name", SEIRMS_CASES)
def test_seirms_progression_matches_reference(name):
    g = load_interventions_golden(name)
    port = SeirmsPort(g["age"], g["stage0"], g["edges"],
                      recovered_to_susceptible_time=g["overrides"]["meta"]["RECOVERED_TO_SUSCEPTIBLE_TIME"])
    out = seirms_episode_with_grads(port, g["seed"], g["mode"], float(g["p_isolate"]), g["num_steps"], g["r2"], g["m"],
                                    g["gI"], g["gD"])
    assert np.array_equal(out["disease_stage_history"], g["ref_disease_stage_history"].astype(np.float64))
    assert np.array_equal(out["next_stage_time_history"], g["ref_next_stage_time_history"].astype(np.float64))
    assert rel_err(g["ref_daily_infected"], out["daily_infected"]) <= 1e-6
    assert rel_err(g["ref_daily_deaths"], out["daily_deaths"]) <= 1e-6
    assert rel_err(g["ref_grad_R2"], out["grad_R2"]) <= 1e-6
    assert rel_err(g["ref_grad_M"], out["grad_M"]) <= 1e-6
    assert abs(out["loss"] - float(g["ref_loss"])) <= 1e-6 * abs(float(g["ref_loss"]))


def _astoria(name):
    g = load_interventions_golden(name)
    src = np.load(os.path.join(GOLDEN_DIR, "astoria_inputs.npz"))
    for k in ("age", "stage0", "edges"):
        g[k] = src[k]
    return g


def test_bundled_astoria_population_with_interventions():
    """37 518 agents, 136 378 edges (the reference's own population): states at steps 0, 10, 20 bit for bit."""
    g = _astoria("interventions_astoria")
    ov = g["overrides"]
    out = interventions_episode(SeirmPort(g["age"], g["stage0"], g["edges"]), g["seed"], g["mode"], float(g["p_isolate"]),
                                g["num_steps"], g["r2"], g["m"], ov["meta"], ov["initial"], ov["env"])
    steps = list(g["kept_steps"])
    for k in ("disease_stage", "next_stage_time", "is_quarantined", "quarantine_start_date", "awaiting_test_result",
              "test_result_date", "test_re_eligble_date"):
        assert np.array_equal(out[k + "_history"][steps], g["ref_" + k + "_history"].astype(np.float64)), k
    assert rel_err(g["ref_daily_infected"], out["daily_infected"]) <= 1e-6
    assert rel_err(g["ref_daily_deaths"], out["daily_deaths"]) <= 1e-6


def test_bundled_astoria_population_with_seirms_progression():
    g = _astoria("seirms_astoria")
    port = SeirmsPort(g["age"], g["stage0"], g["edges"],
                      recovered_to_susceptible_time=g["overrides"]["meta"]["RECOVERED_TO_SUSCEPTIBLE_TIME"])
    out = seirms_episode_with_grads(port, g["seed"], g["mode"], float(g["p_isolate"]), g["num_steps"], g["r2"], g["m"],
                                    g["gI"], g["gD"])
    steps = list(g["kept_steps"])
    assert np.array_equal(out["disease_stage_history"][steps], g["ref_disease_stage_history"].astype(np.float64))
    assert np.array_equal(out["next_stage_time_history"][steps], g["ref_next_stage_time_history"].astype(np.float64))
    assert rel_err(g["ref_daily_deaths"], out["daily_deaths"]) <= 1e-6
    assert rel_err(g["ref_grad_R2"], out["grad_R2"]) <= 1e-6
    assert rel_err(g["ref_grad_M"], out["grad_M"]) <= 1e-6
