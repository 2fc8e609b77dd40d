"""GPU parity through the reference-facing API: envs.create / runner.step / loss.backward, both the generic
Controller path (per-substep custom ops) and the fused episode path, against the real reference's goldens."""
import numpy as np
import pytest
import torch

from tests.helpers import load_golden, make_noise, rel_err

pytestmark = pytest.mark.gpu


def _make_runner(g, tmp_path, mode, fused):
    from agenttorch_b200.core import envs
    from agenttorch_b200.models import covid
    from agenttorch_b200.populations import write_population_folder

    folder = write_population_folder(str(tmp_path / "pop"), g["age"], g["stage0"], g["edges"])
    runner = envs.create(model=covid, population=folder)
    runner._fused_enabled = fused
    tr = runner.initializer.transition_function["0"]["new_transmission"]
    pg = runner.initializer.transition_function["1"]["seirm_progression"]
    pol = runner.initializer.policy_function["0"]["citizens"]["make_isolation_decision"]
    dev = runner.initializer.device
    assert dev.type == "cuda"
    r2 = torch.tensor(g["r2"], dtype=torch.float32, device=dev, requires_grad=True)
    m = torch.tensor(g["m"], dtype=torch.float32, device=dev, requires_grad=True)
    runner._set_parameters({"initializer.transition_function.0.new_transmission.learnable_args.R2": r2,
                            "initializer.transition_function.1.seirm_progression.learnable_args.M": m})
    T = g["num_steps"]
    if mode == "inject":
        U, W = make_noise(g["seed"], g["mode"], len(g["age"]), T, float(g["p_isolate"]))
        Ut, Wt = torch.from_numpy(U).to(dev), torch.from_numpy(W).to(dev)
        pol.mode = "injected"
        pol.behavior = lambda obs: Wt[int(runner.state["current_step"])].reshape(-1, 1)
        tr.exposure_uniform = lambda t: Ut[t]
    else:
        tr.noise_seed, tr.advance_seed = g["seed"], False
    return runner, r2, m


def _check(g, runner, r2, m):
    T = g["num_steps"]
    env = runner.state_trajectory[-1][-1]["environment"]
    di, dd = env["daily_infected"], env["daily_deaths"]
    dev = di.device
    loss = (torch.from_numpy(g["gI"]).to(dev) * di.reshape(-1)).sum() + (torch.from_numpy(g["gD"]).to(dev) * dd.reshape(-1)).sum()
    loss.backward()
    ag = runner.state["agents"]["citizens"]
    assert np.array_equal(ag["disease_stage"].detach().cpu().numpy().reshape(-1).astype(np.float64), g["ref_stage"])
    assert np.array_equal(ag["next_stage_time"].cpu().numpy().reshape(-1).astype(np.int64), g["ref_next_stage_time"].astype(np.int64))
    assert np.array_equal(ag["infected_time"].cpu().numpy().reshape(-1).astype(np.int64), g["ref_infected_time"].astype(np.int64))
    assert tuple(dd.shape) == (1, T)                                   # Q6
    assert rel_err(g["ref_daily_infected"], di.detach().cpu().numpy()) <= 1e-5
    assert rel_err(g["ref_daily_deaths"], dd.detach().cpu().numpy()) <= 1e-5
    assert rel_err(g["ref_grad_R2"], r2.grad.cpu().numpy()) <= 1e-4
    assert rel_err(g["ref_grad_M"], m.grad.cpu().numpy()) <= 1e-4
    assert abs(float(loss) - float(g["ref_loss"])) <= 1e-5 * abs(float(g["ref_loss"]))


@pytest.mark.parametrize("case", ["synth3k_hard", "synth3k_soft", "synth3k_r2vec", "tiny96_hard", "astoria_hard"])
def test_generic_path_with_injected_noise(case, tmp_path):
    g = load_golden(case)
    runner, r2, m = _make_runner(g, tmp_path, "inject", fused=False)
    runner.step(g["num_steps"])
    assert runner.perf_stats["fused_episodes"] == 0
    assert len(runner.state_trajectory) == g["num_steps"] + 1 and len(runner.state_trajectory[-1]) == 2
    _check(g, runner, r2, m)


@pytest.mark.parametrize("case", ["synth3k_soft", "astoria_soft"])
def test_fused_episode_path(case, tmp_path):
    g = load_golden(case)
    runner, r2, m = _make_runner(g, tmp_path, "hash", fused=True)
    runner.step(g["num_steps"])
    assert runner.perf_stats["fused_episodes"] == 1 and runner.perf_stats["kernel_launches"] >= g["num_steps"]
    mid = runner.state_trajectory[5][0]["environment"]["daily_infected"]   # after substep '0' of step 4
    assert rel_err(g["ref_daily_infected"][:5], mid.detach().cpu().numpy()[:5]) <= 1e-5 and float(mid[5:].abs().sum()) == 0.0
    _check(g, runner, r2, m)


def test_fused_and_generic_paths_agree_and_reset_state(tmp_path):
    g = load_golden("synth3k_soft")
    runner, r2, m = _make_runner(g, tmp_path, "hash", fused=True)
    runner.step(g["num_steps"])
    di1 = runner.state["environment"]["daily_infected"].detach().clone()
    runner.reset_state()                         # keeps modules / parameters, re-reads the state
    runner.step(g["num_steps"])                  # CUDA-graph replay on new state tensors
    assert torch.equal(di1, runner.state["environment"]["daily_infected"].detach())
    _check(g, runner, r2, m)
    # the default policy draws torch.rand: just check the loop runs and produces a plausible epidemic
    runner2, _, _ = _make_runner(g, tmp_path, "hash", fused=False)
    runner2.step(3)
    assert float(runner2.state["environment"]["daily_infected"].sum()) >= 0.0


def test_fused_episode_learnable_tables(tmp_path):
    """SFSusceptibility / SFInfector as differentiable environment tensors: loss.backward() through runner.step() fills their
    gradients as the reference's autograd does (golden from the real reference, oracle/ref_harness/make_golden.py round2)"""
    g = load_golden("synth3k_tables_soft")
    runner, r2, m = _make_runner(g, tmp_path, "hash", fused=True)
    env = runner.state["environment"]
    env["SFSusceptibility"] = env["SFSusceptibility"].detach().clone().requires_grad_(True)
    env["SFInfector"] = env["SFInfector"].detach().clone().requires_grad_(True)
    runner.step(g["num_steps"])
    assert runner.perf_stats["fused_episodes"] == 1
    _check(g, runner, r2, m)
    assert rel_err(g["ref_grad_SFS"], env["SFSusceptibility"].grad.cpu().numpy()) <= 1e-4
    assert rel_err(g["ref_grad_SFI"][1:3], env["SFInfector"].grad.cpu().numpy().reshape(-1)[1:3]) <= 1e-4


def test_repeated_fused_episodes_with_restore_state(tmp_path):
    """the same runner, restored and stepped again with other parameters (CUDA-graph replay; the 'epidemic over' short cuts leave
    dne_dr and the code buffers of late steps untouched): every episode must come out as on a fresh runner"""
    from agenttorch_b200.core.snapshot import PackedSnapshot

    g_soft = load_golden("synth3k_soft")
    runner, r2, m = _make_runner(g_soft, tmp_path, "hash", fused=True)
    cit, env = runner.state["agents"]["citizens"], runner.state["environment"]
    snap = PackedSnapshot({"agents": {"citizens": {k: cit[k].clone() for k in ("disease_stage", "next_stage_time", "infected_time")}},
                           "environment": {"daily_infected": env["daily_infected"].clone(), "daily_deaths": env["daily_deaths"].clone()}},
                          compact={"disease_stage": torch.uint8, "next_stage_time": torch.int16, "infected_time": torch.int16})
    dev = runner.initializer.device
    for rep in range(5):
        runner.restore_state(snap)
        if rep >= 2:
            snap.prefetch(dev)                                       # episodes 3.. restore from the prefetched staging buffer
        if rep == 1:                                                 # a very different episode in between: R2 x 6
            big = torch.tensor(g_soft["r2"] * 6, dtype=torch.float32, device=dev, requires_grad=True)
            runner._set_parameters({"initializer.transition_function.0.new_transmission.learnable_args.R2": big})
            runner.step(g_soft["num_steps"])
            runner.state["environment"]["daily_infected"].sum().backward()
            continue
        r2 = torch.tensor(g_soft["r2"], dtype=torch.float32, device=dev, requires_grad=True)
        m = torch.tensor(g_soft["m"], dtype=torch.float32, device=dev, requires_grad=True)
        runner._set_parameters({"initializer.transition_function.0.new_transmission.learnable_args.R2": r2,
                                "initializer.transition_function.1.seirm_progression.learnable_args.M": m})
        runner.step(g_soft["num_steps"])
        _check(g_soft, runner, r2, m)


@pytest.mark.gpu
def test_r_steps_on_the_device_equals_the_per_step_expression():
    """the fused per-step R (one broadcast product + row sum) must reproduce the reference's per-step `(R2 * one_hot[week]).sum()`
    bit for bit on the device too (R enters the exposure probabilities, so a different rounding would change trajectories)"""
    dev = torch.device("cuda", 0)
    eye = torch.eye(3, device=dev)
    gen = torch.Generator(device="cpu").manual_seed(11)
    for trial in range(200):
        for shape in ((3, 1), (3,)):
            r = (torch.rand(shape, generator=gen) * 9 + 0.1).to(dev)
            per_week = torch.stack([(r * eye[w]).sum() for w in range(3)])
            fused = (r.unsqueeze(0) * eye.view(3, *([1] * (r.dim() - 1)), 3)).flatten(1).sum(1)
            assert torch.equal(per_week, fused)
