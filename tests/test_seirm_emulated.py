"""The SEIRM engine (agenttorch_b200/engine.py) driven against csrc/at_seirm.cu COMPILED FOR THE HOST on a small SIMT
emulator (tests/host_emulation/simt_shim.h: one OS thread per CUDA thread, block/warp rendezvous for __syncthreads and the
warp intrinsics, TMA bulk copies as memcpy + mbarrier phase counts): the step kernel in its three layouts, the update-only kernel, the
filter fold and the fused backward kernel are checked against the real reference's golden outputs on a machine without a
GPU.  The -m gpu tests (tests/test_gpu_parity.py) run the same comparisons on the real library."""
import contextlib
import os
import subprocess
import types

import numpy as np
import pytest
import torch

from agenttorch_b200 import _native, engine
from tests.helpers import load_golden, make_noise, rel_err

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def emulated_library(tmp_path_factory):
    out = str(tmp_path_factory.mktemp("seirm_emu") / "libat_seirm_emulated.so")
    cmd = ["g++", "-O1", "-std=c++17", "-ffp-contract=off", "-shared", "-fPIC", "-pthread", "-I", os.path.join(ROOT, "include"),
           os.path.join(ROOT, "tests", "host_emulation", "seirm_emulation.cpp"), "-o", out]
    res = subprocess.run(cmd, capture_output=True, text=True)
    assert res.returncode == 0, res.stderr
    return out


@pytest.fixture()
def engine_on_host(emulated_library, monkeypatch):
    """Point the ctypes binding at the emulated library; lift the engine's CUDA checks and stream / device calls."""
    monkeypatch.setattr(_native, "LIB_PATH", emulated_library)
    monkeypatch.setattr(_native, "_lib", None)
    monkeypatch.setattr(engine, "_require_cuda", lambda message, is_cuda: None)
    monkeypatch.setattr(torch.cuda, "device", lambda device: contextlib.nullcontext())
    monkeypatch.setattr(torch.cuda, "current_stream", lambda device=None: types.SimpleNamespace(cuda_stream=0))
    monkeypatch.setattr(torch.cuda, "synchronize", lambda device=None: None)
    yield
    monkeypatch.setattr(_native, "_lib", None)              # the next user loads the real library again


def _episode(g, *, packed=True, tensor_noise=False, before=None, **engine_kw):
    from agenttorch_b200.engine import NoiseSpec, SeirmEngine, SeirmSpec, SeirmState
    from agenttorch_b200.graph import build_contact_graph, mean_interactions_from_age

    T, n = g["num_steps"], len(g["age"])
    age = torch.from_numpy(g["age"].astype(np.int64))
    edges = torch.from_numpy(g["edges"].astype(np.int64)).t().contiguous()
    graph = build_contact_graph(edges, n, mean_interactions=mean_interactions_from_age(age))
    engine_kw.setdefault("warps_per_cta", 2)                                                # 64 OS threads per block
    eng = SeirmEngine(graph, SeirmSpec(num_steps=T), **engine_kw)
    st = SeirmState.from_initial(g["stage0"].astype(np.float32), g["age"], T, torch.device("cpu"), packed=packed)
    if tensor_noise:
        U, W = make_noise(g["seed"], g["mode"], n, T, float(g["p_isolate"]))
        noise = NoiseSpec(seed=0, isolation="tensor", uniform="tensor", isolation_tensor=torch.from_numpy(W),
                          uniform_tensor=torch.from_numpy(U))
    else:
        noise = NoiseSpec(seed=g["seed"], isolation=g["mode"], p_isolate=float(g["p_isolate"]))
    r2 = torch.tensor(g["r2"], dtype=torch.float32, requires_grad=True)
    eye = torch.eye(3)
    r_steps = torch.stack([(r2 * eye[int(t / 7)]).sum() for t in range(T)])
    m = torch.tensor(g["m"], dtype=torch.float32)
    if before is not None:
        # another episode on the SAME engine first (same population, other parameters / isolation mode): its saved-for-backward
        # buffers, code buffers and "epidemic over" state must not leak into the episode that is checked
        st_b = SeirmState.from_initial(before["stage0"].astype(np.float32), before["age"], T, torch.device("cpu"), packed=packed)
        rb = torch.stack([(torch.tensor(before["r2"], dtype=torch.float32) * eye[int(t / 7)]).sum() for t in range(T)])
        eng.run_episode(st_b, rb, torch.tensor(before["m"], dtype=torch.float32), T,
                        NoiseSpec(seed=before["seed"], isolation=before["mode"], p_isolate=float(before["p_isolate"])), want_grad=True)
    sv = eng.run_episode(st, r_steps.detach(), m, T, noise, want_grad=True, cuda_graph=False)
    gr, gm = eng.episode_backward(sv, T, torch.from_numpy(g["gI"]), torch.from_numpy(g["gD"]), m)
    eng.check_errors()
    r_steps.backward(gr)
    return st, r2.grad.numpy(), gm.numpy()


def _check(g, st, grad_r2, grad_m):
    assert np.array_equal(st.stage.numpy().astype(np.float64), g["ref_stage"])                                 # bit-exact
    assert np.array_equal(st.next_stage_time.numpy().astype(np.int64), g["ref_next_stage_time"].astype(np.int64))
    assert np.array_equal(st.infected_time.numpy().astype(np.int64), g["ref_infected_time"].astype(np.int64))
    assert rel_err(g["ref_daily_infected"], st.daily_infected.numpy()) <= 1e-5
    assert rel_err(g["ref_daily_deaths"], st.daily_deaths.numpy()) <= 1e-5
    assert rel_err(g["ref_grad_R2"], grad_r2) <= 1e-4
    assert rel_err(g["ref_grad_M"], grad_m) <= 1e-4


@pytest.mark.parametrize("case,kw", [
    ("tiny96_hard", dict(layout="bcsr")),                            # fused kernel, exact shared-memory bitmap
    ("tiny96_hard", dict(layout="bcsr", packed=False, tensor_noise=True)),   # float32 timers, injected uniforms
    ("noiso500", dict(layout="bcsr")),
    ("synth3k_hard", dict(layout="bcsr")),
    ("synth3k_soft", dict(layout="bcsr", block_size=1024)),          # column blocks: gather + update kernels
    ("synth3k_r2vec", dict(hashed=True, hash_words=32)),             # hashed filter + hot_fold_kernel
    ("tiny96_hard", dict()),                                         # stream layout (the default): one block, one tile
    ("noiso500", dict()),
    ("synth3k_hard", dict()),
    ("synth3k_soft", dict(block_size=1024, tile_rows=256)),          # stream layout: 3 source blocks, 12 tiles, CTAs spanning blocks
    ("synth3k_hard", dict(block_size=2048, tile_rows=2048, grid=5)),
    ("synth3k_r2vec", dict(block_size=4096, tile_rows=64, warps_per_cta=1)),   # one block, tiny tiles
    ("tiny96_hard", dict(block_size=1024, tile_rows=32, grid=1)),
    ("noiso500", dict(block_size=1024, tile_rows=128, packed=False, tensor_noise=True)),
    ("tiny96_hard", dict(layout="bcsr", warps_per_cta=1)),           # a CTA smaller than the code tables (found by tools/fuzz_emulated.py)
    ("noiso500", dict(warps_per_cta=1, hashed=True, hash_words=32)),
])
def test_kernel_source_matches_reference_golden(engine_on_host, case, kw):
    g = load_golden(case)
    st, grad_r2, grad_m = _episode(g, **kw)
    _check(g, st, grad_r2, grad_m)


@pytest.mark.parametrize("case,kw", [
    ("layers400_none", dict(block_size=1024, tile_rows=64)),          # three network layers, one / several source blocks
    ("synth3k_layers_hard", dict(block_size=1024, tile_rows=256)),
    ("synth3k_layers_hard", dict()),
    ("synth3k_tables_soft", dict()),
])
def test_layers_and_table_gradients_match_reference_golden(engine_on_host, case, kw):
    """multi-layer contact networks (per-edge network type -> per-edge I_bar, transition.py:71-73) and the gradients with respect
    to SFSusceptibility / SFInfector, against outputs of the real reference (oracle/ref_harness/make_golden.py round2)"""
    from agenttorch_b200.engine import NoiseSpec, SeirmEngine, SeirmSpec, SeirmState
    from agenttorch_b200.graph import build_contact_graph, mean_interactions_from_age

    g = load_golden(case)
    T, n = g["num_steps"], len(g["age"])
    age = torch.from_numpy(g["age"].astype(np.int64))
    edges = torch.from_numpy(g["edges"].astype(np.int64)).t().contiguous()
    attr = None
    if "edge_type" in g:
        attr = torch.stack((torch.from_numpy(g["edge_type"]).float(), torch.ones(edges.shape[1])))
    graph = build_contact_graph(edges, n, edge_attr=attr, mean_interactions=mean_interactions_from_age(age))
    assert (len(graph.i_bar_layers) == 3) == ("edge_type" in g)
    eng = SeirmEngine(graph, SeirmSpec(num_steps=T), warps_per_cta=2, **kw)
    eng.want_table_grad = True
    st = SeirmState.from_initial(g["stage0"].astype(np.float32), g["age"], T, torch.device("cpu"))
    noise = NoiseSpec(seed=g["seed"], isolation=g["mode"], p_isolate=float(g["p_isolate"]))
    r2 = torch.tensor(g["r2"], dtype=torch.float32, requires_grad=True)
    eye = torch.eye(3)
    r_steps = torch.stack([(r2 * eye[int(t / 7)]).sum() for t in range(T)])
    m = torch.tensor(g["m"], dtype=torch.float32)
    sv = eng.run_episode(st, r_steps.detach(), m, T, noise, want_grad=True)
    gI, gD = torch.from_numpy(g["gI"]), torch.from_numpy(g["gD"])
    gr, gm = eng.episode_backward(sv, T, gI, gD, m)
    r_steps.backward(gr)
    _check(g, st, r2.grad.numpy(), gm.numpy())
    g_sfs, g_sfi = eng.table_gradients(sv, T, gI, gD, m, r_steps.detach(), st.age)
    assert rel_err(g["ref_grad_SFS"], g_sfs.numpy()) <= 1e-4
    # SFInfector[1] (exposed) and [2] (infected) sources; [0] — fractional-stage sources under soft isolation — is not tracked
    assert rel_err(g["ref_grad_SFI"][1:3], g_sfi.numpy()[1:3]) <= 1e-4
    if g["mode"] != "soft":
        assert rel_err(g["ref_grad_SFI"], g_sfi.numpy()) <= 1e-4


@pytest.mark.parametrize("first,second,kw", [
    ("synth3k_hard", "synth3k_soft", dict()),                        # a spreading epidemic, then one that is over after 9 steps
    ("synth3k_soft", "synth3k_hard", dict(block_size=1024, tile_rows=256)),
    ("synth3k_hard", "synth3k_soft", dict(layout="bcsr")),
])
def test_consecutive_episodes_do_not_leak_into_each_other(engine_on_host, first, second, kw):
    """Once nobody is in E or I the stream layout skips the gather, the zero-fill of dne_dr and (one step later) of the code
    buffers; whatever an EARLIER episode left in those buffers must then not show up in the results or the gradients."""
    g = load_golden(second)
    st, grad_r2, grad_m = _episode(g, before=load_golden(first), **kw)
    _check(g, st, grad_r2, grad_m)


def test_the_real_binding_is_restored_after_emulation():
    assert _native.LIB_PATH.endswith(os.path.join("lib", "libat_seirm.so"))
    assert engine._require_cuda.__module__ == "agenttorch_b200.engine"
    from agenttorch_b200.engine import SeirmEngine, SeirmSpec
    from agenttorch_b200.graph import synthetic_contact_graph

    with pytest.raises(RuntimeError, match="CUDA"):
        SeirmEngine(synthetic_contact_graph(512, 4.0, 1, device="cpu"), SeirmSpec())


# ------------------------------------------------------------------------------------------------------------------
# The reference-facing API (envs.create / runner.step / loss.backward) on the emulated library: the generic Controller path
# (one custom op per substep, registered autograd) and the fused-episode path, as in tests/test_gpu_runner.py.
def _runner(g, tmp_path, monkeypatch, fused):
    from agenttorch_b200.core import envs
    from agenttorch_b200.models import covid
    from agenttorch_b200.models.covid import context
    from agenttorch_b200.populations import write_population_folder

    monkeypatch.setattr(context, "_require_cuda", lambda is_cuda: None)
    monkeypatch.setattr(context, "_CACHE", {})
    monkeypatch.setattr(context, "SeirmEngine", lambda graph, spec: engine.SeirmEngine(graph, spec, warps_per_cta=2))
    folder = write_population_folder(str(tmp_path / "pop"), g["age"], g["stage0"], g["edges"])
    runner = envs.create(model=covid, population=folder)
    assert runner.initializer.device.type == "cpu"
    runner._fused_enabled, runner._use_cuda_graph = fused, False
    tr = runner.initializer.transition_function["0"]["new_transmission"]
    pg = runner.initializer.transition_function["1"]["seirm_progression"]
    pol = runner.initializer.policy_function["0"]["citizens"]["make_isolation_decision"]
    r2 = torch.tensor(g["r2"], dtype=torch.float32, requires_grad=True)
    m = torch.tensor(g["m"], dtype=torch.float32, requires_grad=True)
    runner._set_parameters({"initializer.transition_function.0.new_transmission.learnable_args.R2": r2,
                            "initializer.transition_function.1.seirm_progression.learnable_args.M": m})
    if fused:
        tr.noise_seed, tr.advance_seed = g["seed"], False
        runner._plan_cache = {"transmission": tr, "progression": pg, "policy": pol}     # the planner insists on a CUDA device
    else:
        U, W = make_noise(g["seed"], g["mode"], len(g["age"]), g["num_steps"], float(g["p_isolate"]))
        Ut, Wt = torch.from_numpy(U), torch.from_numpy(W)
        pol.mode = "injected"
        pol.behavior = lambda obs: Wt[int(runner.state["current_step"])].reshape(-1, 1)
        tr.exposure_uniform = lambda t: Ut[t]
    return runner, r2, m


def _check_runner(g, runner, r2, m):
    env = runner.state_trajectory[-1][-1]["environment"]
    di, dd = env["daily_infected"], env["daily_deaths"]
    loss = (torch.from_numpy(g["gI"]) * di.reshape(-1)).sum() + (torch.from_numpy(g["gD"]) * dd.reshape(-1)).sum()
    loss.backward()
    ag = runner.state["agents"]["citizens"]
    assert np.array_equal(ag["disease_stage"].detach().numpy().reshape(-1).astype(np.float64), g["ref_stage"])
    assert np.array_equal(ag["next_stage_time"].numpy().reshape(-1).astype(np.int64), g["ref_next_stage_time"].astype(np.int64))
    assert np.array_equal(ag["infected_time"].numpy().reshape(-1).astype(np.int64), g["ref_infected_time"].astype(np.int64))
    assert tuple(dd.shape) == (1, g["num_steps"])
    assert rel_err(g["ref_daily_infected"], di.detach().numpy()) <= 1e-5
    assert rel_err(g["ref_daily_deaths"], dd.detach().numpy()) <= 1e-5
    assert rel_err(g["ref_grad_R2"], r2.grad.numpy()) <= 1e-4
    assert rel_err(g["ref_grad_M"], m.grad.numpy()) <= 1e-4
    assert abs(float(loss) - float(g["ref_loss"])) <= 1e-5 * abs(float(g["ref_loss"]))


@pytest.mark.parametrize("case", ["tiny96_hard", "synth3k_r2vec"])
def test_runner_generic_path_on_host(engine_on_host, monkeypatch, tmp_path, case):
    g = load_golden(case)
    runner, r2, m = _runner(g, tmp_path, monkeypatch, fused=False)
    runner.step(g["num_steps"])
    assert runner.perf_stats["fused_episodes"] == 0 and len(runner.state_trajectory) == g["num_steps"] + 1
    _check_runner(g, runner, r2, m)


def test_runner_fused_episode_path_on_host(engine_on_host, monkeypatch, tmp_path):
    g = load_golden("synth3k_soft")
    runner, r2, m = _runner(g, tmp_path, monkeypatch, fused=True)
    runner.step(g["num_steps"])
    assert runner.perf_stats["fused_episodes"] == 1
    _check_runner(g, runner, r2, m)


def test_runner_fused_episode_learnable_tables_on_host(engine_on_host, monkeypatch, tmp_path):
    """SFSusceptibility / SFInfector as differentiable environment tensors (learnable-capable in config.yaml): loss.backward()
    through runner.step() fills their gradients, as in the reference (golden from the real reference)"""
    g = load_golden("synth3k_tables_soft")
    runner, r2, m = _runner(g, tmp_path, monkeypatch, fused=True)
    env = runner.state["environment"]
    env["SFSusceptibility"] = env["SFSusceptibility"].detach().clone().requires_grad_(True)
    env["SFInfector"] = env["SFInfector"].detach().clone().requires_grad_(True)
    runner.step(g["num_steps"])
    assert runner.perf_stats["fused_episodes"] == 1
    _check_runner(g, runner, r2, m)
    assert rel_err(g["ref_grad_SFS"], env["SFSusceptibility"].grad.numpy()) <= 1e-4
    assert rel_err(g["ref_grad_SFI"][1:3], env["SFInfector"].grad.numpy().reshape(-1)[1:3]) <= 1e-4


# ------------------------------------------------------------------------------------------------------------------
# Fused episode of a config WITH the quarantine / testing substeps: the SEIRM step launches and, after each, one in-place
# launch per intervention block (models/covid/interventions.py::plan_intervention_substeps), both libraries emulated.
@pytest.fixture(scope="module")
def emulated_interventions_library(tmp_path_factory):
    out = str(tmp_path_factory.mktemp("iv_emu") / "libat_interventions_emulated.so")
    cmd = ["g++", "-O1", "-std=c++17", "-shared", "-fPIC", "-I", os.path.join(ROOT, "include"),
           os.path.join(ROOT, "tests", "host_emulation", "iv_emulation.cpp"), "-o", out]
    res = subprocess.run(cmd, capture_output=True, text=True)
    assert res.returncode == 0, res.stderr
    return out


@pytest.mark.parametrize("name", ["interventions_1500_softopen", "interventions_1500_default"])
def test_fused_episode_with_intervention_substeps_on_host(engine_on_host, emulated_interventions_library, monkeypatch, tmp_path, name):
    from agenttorch_b200 import _native_iv
    from agenttorch_b200 import interventions as iv_ops
    from agenttorch_b200.core.dataloader import DataLoader, LoadPopulation
    from agenttorch_b200.core.runner import Runner
    from agenttorch_b200.models import covid
    from agenttorch_b200.models.covid import context
    from agenttorch_b200.models.covid.interventions import add_intervention_substeps, plan_intervention_substeps
    from agenttorch_b200.populations import write_population_folder
    from tests.helpers import load_interventions_golden

    monkeypatch.setattr(_native_iv, "LIB_PATH", emulated_interventions_library)
    monkeypatch.setattr(_native_iv, "_lib", None)
    monkeypatch.setattr(iv_ops, "_require_cuda", lambda name, t: None)
    monkeypatch.setattr(iv_ops, "_stream", lambda t: None)
    monkeypatch.setattr(context, "_require_cuda", lambda is_cuda: None)
    monkeypatch.setattr(context, "_CACHE", {})
    monkeypatch.setattr(context, "SeirmEngine", lambda graph, spec: engine.SeirmEngine(graph, spec, warps_per_cta=2))

    g = load_interventions_golden(name)
    ov = g["overrides"]
    folder = write_population_folder(str(tmp_path / "pop"), g["age"], g["stage0"], g["edges"])
    loader = DataLoader(covid, LoadPopulation(folder))
    for k, v in ov["meta"].items():
        loader.set_config_attribute(k, v)
    loader.set_config_attribute("noise_seed", g["seed"])
    for k, v in ov["initial"].items():
        loader.config["state"]["agents"]["citizens"]["properties"][k]["value"] = v
    for k, v in ov["env"].items():
        loader.config["state"]["environment"][k]["value"] = v
    runner = Runner(add_intervention_substeps(loader.get_config()), covid.get_registry())
    runner.init()
    runner._use_cuda_graph = False
    tf, pf = runner.initializer.transition_function, runner.initializer.policy_function
    tr, pg = tf["0"]["new_transmission"], tf["1"]["seirm_progression"]
    noisy = [tr, pf["2"]["citizens"]["start_compliance"], pf["2"]["citizens"]["break_compliance"], pf["3"]["citizens"]["accept_test"]]
    for mod in noisy:
        assert mod.noise_seed == g["seed"]
        mod.advance_seed = False
    r2 = torch.tensor(g["r2"], dtype=torch.float32, requires_grad=True)
    m = torch.tensor(g["m"], dtype=torch.float32, requires_grad=True)
    runner._set_parameters({"initializer.transition_function.0.new_transmission.learnable_args.R2": r2,
                            "initializer.transition_function.1.seirm_progression.learnable_args.M": m})
    assert runner._fused_plan() is None                                   # the planner insists on a CUDA device ...
    runner._plan_cache = None
    real_device = runner.initializer.device
    runner.initializer.device = types.SimpleNamespace(type="cuda")        # ... so it is shown one for the duration of the planning
    plan = runner._fused_plan()
    runner.initializer.device = real_device
    assert plan is not None and plan["after_step"] is not None and plan["transmission"] is tr and plan["progression"] is pg
    T = g["num_steps"]
    runner.step(T)
    assert runner.perf_stats["fused_episodes"] == 1 and runner.perf_stats["kernel_launches"] >= 3 * T
    assert len(runner.state_trajectory) == T + 1 and len(runner.state_trajectory[-1]) == 4
    cit = runner.state["agents"]["citizens"]
    for k in ("disease_stage", "next_stage_time", "is_quarantined", "quarantine_start_date", "awaiting_test_result",
              "test_result_date", "test_re_eligble_date"):
        got = cit[k].detach().numpy().reshape(-1).astype(np.float64)
        assert np.array_equal(got, g["ref_" + k + "_history"][T - 1].astype(np.float64)), k
    env = runner.state["environment"]
    di, dd = env["daily_infected"], env["daily_deaths"]
    assert rel_err(g["ref_daily_infected"], di.detach().numpy()) <= 1e-5 and rel_err(g["ref_daily_deaths"], dd.detach().numpy()) <= 1e-5
    loss = (torch.from_numpy(g["gI"]) * di.reshape(-1)).sum() + (torch.from_numpy(g["gD"]) * dd.reshape(-1)).sum()
    loss.backward()
    assert rel_err(g["ref_grad_R2"], r2.grad.numpy()) <= 1e-4 and rel_err(g["ref_grad_M"], m.grad.numpy()) <= 1e-4
    # a second episode replays the cached launch plan: the hook must fire there as well
    before = cit["is_quarantined"].clone()
    runner.step(2)
    assert not torch.equal(before, runner.state["agents"]["citizens"]["is_quarantined"]) or name.endswith("default")
    # injected draws or policies with different noise addresses are not fused — also when that changes after planning
    assert runner._plan_still_valid(plan)
    pf["3"]["citizens"]["accept_test"].uniform = lambda t: None
    assert plan_intervention_substeps(runner, ["2", "3"]) is None and not runner._plan_still_valid(plan)
    pf["3"]["citizens"]["accept_test"].uniform = None
    pf["2"]["citizens"]["break_compliance"].noise_seed += 1
    assert plan_intervention_substeps(runner, ["2", "3"]) is None and not runner._plan_still_valid(plan)


# ------------------------------------------------------------------------------------------------------------------
# Seeded synthetic inputs against the oracle, as in tests/test_gpu_synthetic.py: tasks spanning several TMA chunks, weighted
# edges, an empty graph — small enough for the emulator.
def _synthetic(n, deg, seed, frac_e=0.02, frac_i=0.01):
    rng = np.random.default_rng(seed)
    age = rng.integers(0, 6, n)
    stage0 = np.zeros(n)
    r = rng.random(n)
    stage0[r < frac_e] = 1.0
    stage0[(r >= frac_e) & (r < frac_e + frac_i)] = 2.0
    edges = np.stack([rng.integers(0, n, int(n * deg)), rng.integers(0, n, int(n * deg))], axis=1)
    return age, stage0, edges


def _against_oracle(age, stage0, edges, seed, mode, r2, m, weights=None, **engine_kw):
    from agenttorch_b200.engine import NoiseSpec, SeirmEngine, SeirmSpec, SeirmState
    from agenttorch_b200.graph import build_contact_graph, mean_interactions_from_age
    from oracle.seirm_port import SeirmPort, episode_with_grads

    T, n = 21, len(age)
    U, W = make_noise(seed, mode, n, T)
    gI, gD = np.linspace(0.5, 1.5, T, dtype=np.float32), np.linspace(2.0, 1.0, T, dtype=np.float32)
    ref = episode_with_grads(SeirmPort(age, stage0, edges, edge_weight=weights), W, U, gI, gD, r2=r2, m=m)
    ei = torch.from_numpy(edges.astype(np.int64)).t().contiguous()
    attr = None if weights is None else torch.stack((torch.ones(len(edges)), torch.from_numpy(weights.astype(np.float32))))
    graph = build_contact_graph(ei, n, edge_attr=attr, mean_interactions=mean_interactions_from_age(torch.from_numpy(age)))
    eng = SeirmEngine(graph, SeirmSpec(num_steps=T), warps_per_cta=2, **engine_kw)
    st = SeirmState.from_initial(stage0.astype(np.float32), age, T, torch.device("cpu"))
    r2_t = torch.tensor(r2, dtype=torch.float32, requires_grad=True)
    eye = torch.eye(3)
    r_steps = torch.stack([(r2_t * eye[int(t / 7)]).sum() for t in range(T)])
    m_t = torch.tensor(m, dtype=torch.float32)
    sv = eng.run_episode(st, r_steps.detach(), m_t, T, NoiseSpec(seed=seed, isolation=mode), want_grad=True)
    gr, gm = eng.episode_backward(sv, T, torch.from_numpy(gI), torch.from_numpy(gD), m_t)
    eng.check_errors()
    r_steps.backward(gr)
    assert np.array_equal(st.stage.numpy().astype(np.float64), ref["stage"])
    assert np.array_equal(st.next_stage_time.numpy().astype(np.float32), ref["next_stage_time"])
    assert np.array_equal(st.infected_time.numpy().astype(np.float32), ref["infected_time"])
    assert rel_err(ref["daily_infected"], st.daily_infected.numpy()) <= 1e-5 and rel_err(ref["daily_deaths"], st.daily_deaths.numpy()) <= 1e-5
    assert rel_err(ref["grad_R2"], r2_t.grad.numpy()) <= 1e-4 and rel_err(ref["grad_M"], gm.numpy()) <= 1e-4
    return st


_R2, _M = np.array([[0.30], [0.25], [0.20]], dtype=np.float32), np.array([0.2], dtype=np.float32)


@pytest.mark.parametrize("kw", [dict(), dict(edges_per_chunk=512), dict(hash_words=32)])
def test_tasks_larger_than_one_chunk_on_host(engine_on_host, kw):
    """average in-degree 45: a 32-row task holds ~1440 edges = 2-3 chunks of 512/768 edges"""
    age, stage0, edges = _synthetic(1536, 45.0, 3)
    st = _against_oracle(age, stage0, edges, 77, "hard", _R2 * 0.3, _M, **kw)
    assert float(st.daily_infected.sum()) > 20


@pytest.mark.parametrize("kw", [dict(), dict(block_size=1024)])
def test_weighted_edges_on_host(engine_on_host, kw):
    age, stage0, edges = _synthetic(2048, 7.0, 4)
    w = np.random.default_rng(9).uniform(0.25, 2.0, len(edges)).astype(np.float32)
    _against_oracle(age, stage0, edges, 5, "hard", _R2 * 4, _M, weights=w, **kw)


def test_empty_graph_on_host(engine_on_host):
    age, stage0, _ = _synthetic(200, 1.0, 6, frac_e=0.05, frac_i=0.05)
    st = _against_oracle(age, stage0, np.zeros((0, 2), dtype=np.int64), 1, "soft", _R2, _M)
    assert float(st.daily_infected.sum()) == 0.0


def test_after_step_hook_fires_in_recorded_and_replayed_episodes(engine_on_host):
    """SeirmEngine.after_step: called once per step with the step number, both when an episode's launches are recorded and
    when the cached launch plan is replayed; results do not depend on it."""
    from agenttorch_b200.engine import NoiseSpec, SeirmEngine, SeirmSpec, SeirmState
    from agenttorch_b200.graph import build_contact_graph, mean_interactions_from_age

    g = load_golden("tiny96_hard")
    T, n = g["num_steps"], len(g["age"])
    age = torch.from_numpy(g["age"].astype(np.int64))
    graph = build_contact_graph(torch.from_numpy(g["edges"].astype(np.int64)).t().contiguous(), n,
                                mean_interactions=mean_interactions_from_age(age))
    eng = SeirmEngine(graph, SeirmSpec(num_steps=T), warps_per_cta=2)
    fresh = SeirmState.from_initial(g["stage0"].astype(np.float32), g["age"], T, torch.device("cpu"))
    st = fresh.clone()
    noise = NoiseSpec(seed=g["seed"], isolation=g["mode"], p_isolate=float(g["p_isolate"]))
    r_steps, m = torch.full((T,), 4.5), torch.tensor(g["m"], dtype=torch.float32)
    seen, stages = [], []
    eng.after_step = lambda t: (seen.append(t), stages.append(st.stage.clone()))
    results = []
    for _ in range(3):                                       # 1st: recorded; 2nd and 3rd: replayed from the cached plan
        for k in ("stage", "next_stage_time", "infected_time", "daily_infected", "daily_deaths"):
            getattr(st, k).copy_(getattr(fresh, k))
        eng.run_episode(st, r_steps, m, T, noise, want_grad=True, cuda_graph=True)     # a hook forces eager launches
        results.append((st.stage.clone(), st.daily_infected.clone()))
    assert seen == list(range(T)) * 3
    assert all(torch.equal(results[0][0], r[0]) and torch.equal(results[0][1], r[1]) for r in results[1:])
    assert all(torch.equal(stages[t], stages[T + t]) for t in range(T))               # the hook saw the state after step t
    assert torch.equal(stages[T - 1], results[0][0])
    eng.after_step = None


def test_checkpoint_and_resume_reproduce_the_next_episode(engine_on_host, monkeypatch, tmp_path):
    """Runner.save_checkpoint / load_checkpoint: state tensors + parameters + noise counters; an episode run
    from the restored checkpoint is bitwise the episode the original runner runs next."""
    g = load_golden("tiny96_hard")

    def fresh():
        runner, r2, m = _runner(g, tmp_path, monkeypatch, fused=True)
        tr = runner.initializer.transition_function["0"]["new_transmission"]
        tr.noise_seed, tr.advance_seed = 1234, True                      # a new noise stream every episode
        return runner, tr

    a, tr_a = fresh()
    a.step(g["num_steps"])
    path = str(tmp_path / "episode1.ckpt")
    a.save_checkpoint(path)
    assert tr_a.episode_index == 1
    a.step(g["num_steps"])                                                # episode 2 of the original runner
    want = {k: v.detach().clone() for k, v in a.state["agents"]["citizens"].items()}
    want_env = {k: a.state["environment"][k].detach().clone() for k in ("daily_infected", "daily_deaths")}

    b, tr_b = fresh()
    r2_b = b.initializer.transition_function["0"]["new_transmission"].calibrate_R2
    b.load_checkpoint(path)
    assert tr_b.episode_index == 1 and tr_b.noise_seed == 1234
    restored = b.initializer.transition_function["0"]["new_transmission"].calibrate_R2
    assert torch.equal(restored.detach(), torch.tensor(g["r2"])) and restored.requires_grad == r2_b.requires_grad
    b._plan_cache = {"transmission": tr_b, "progression": b.initializer.transition_function["1"]["seirm_progression"],
                     "policy": b.initializer.policy_function["0"]["citizens"]["make_isolation_decision"]}
    b.step(g["num_steps"])
    for k, v in want.items():
        assert torch.equal(b.state["agents"]["citizens"][k].detach(), v), k
    for k, v in want_env.items():
        assert torch.equal(b.state["environment"][k].detach().reshape(-1), v.reshape(-1)), k
    assert float(want_env["daily_infected"].sum()) != float(torch.tensor(g["ref_daily_infected"]).sum())   # a different stream than episode 1
    with pytest.raises(ValueError, match="format"):
        b.load_checkpoint({"format": 99})


def test_infection_older_than_the_table_only_at_the_unused_last_encode(engine_on_host):
    """12-step episode, 12 days of infectiousness: agents exposed at the start are still infectious when the episode ends, and
    their age at step 12 — which is never simulated — is past lam_gamma_integrals.  The reference never looks that entry up; the
    engine used to refuse (the last step encodes the next step's infectiousness).  Found by tools/fuzz_emulated.py."""
    from agenttorch_b200.engine import NoiseSpec, SeirmEngine, SeirmSpec, SeirmState
    from agenttorch_b200.graph import build_contact_graph, mean_interactions_from_age
    from oracle.seirm_port import SeirmConstants, SeirmPort, episode_with_grads

    T = 12
    age, stage0, edges = _synthetic(600, 8.0, 12, frac_e=0.1, frac_i=0.0)
    consts = SeirmConstants(num_steps=T, num_weeks=2, exposed_to_infected_time=5, infected_to_recovered_time=12)
    r2, m = np.array([[1.5], [0.8]], dtype=np.float32), np.array([0.3], dtype=np.float32)
    U, W = make_noise(3, "hard", len(age), T)
    gI, gD = np.linspace(0.5, 1.5, T, dtype=np.float32), np.linspace(2.0, 1.0, T, dtype=np.float32)
    ref = episode_with_grads(SeirmPort(age, stage0, edges, consts=consts), W, U, gI, gD, r2=r2, m=m, num_steps=T)
    graph = build_contact_graph(torch.from_numpy(edges).t().contiguous(), len(age),
                                mean_interactions=mean_interactions_from_age(torch.from_numpy(age)))
    eng = SeirmEngine(graph, SeirmSpec(num_steps=T, infected_to_recovered_time=12), warps_per_cta=2)
    st = SeirmState.from_initial(stage0.astype(np.float32), age, T, torch.device("cpu"))
    r2_t = torch.tensor(r2, requires_grad=True)
    r_steps = torch.stack([(r2_t * torch.eye(2)[int(t / 7)]).sum() for t in range(T)])
    sv = eng.run_episode(st, r_steps.detach(), torch.tensor(m), T, NoiseSpec(seed=3, isolation="hard"), want_grad=True)
    eng.check_errors()                                              # no refusal
    assert np.array_equal(st.stage.numpy().astype(np.float64), ref["stage"]) and float(ref["daily_infected"].sum()) > 20
    assert rel_err(ref["daily_infected"], st.daily_infected.numpy()) <= 1e-5
    # an infection older than the table INSIDE the episode is still refused
    st2 = SeirmState.from_initial(stage0.astype(np.float32), age, T, torch.device("cpu"))
    st2.infected_time.fill_(-9)
    eng.run_episode(st2, r_steps.detach(), torch.tensor(m), T, NoiseSpec(seed=3, isolation="hard"), want_grad=False)
    with pytest.raises(RuntimeError, match="lam_gamma_integrals index out of range"):
        eng.check_errors()
