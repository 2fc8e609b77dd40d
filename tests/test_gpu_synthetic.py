"""GPU parity on seeded synthetic inputs against the CPU oracle (oracle/seirm_port.py), covering what the golden
fixtures do not: tasks larger than one TMA chunk, weighted edges, every engine layout on the same input, bitwise
reproducibility, and size-independent invariants at a size the oracle cannot reach."""
import numpy as np
import pytest
import torch

from oracle.seirm_port import SeirmPort, episode_with_grads
from tests.helpers import make_noise, rel_err

pytestmark = pytest.mark.gpu
T = 21


def _inputs(n, deg, seed, frac_e=0.02, frac_i=0.01):
    rng = np.random.default_rng(seed)
    age = rng.integers(0, 6, n)
    stage0 = np.zeros(n)
    r = rng.random(n)
    stage0[r < frac_e] = 1.0
    stage0[(r >= frac_e) & (r < frac_e + frac_i)] = 2.0
    e = int(n * deg)
    edges = np.stack([rng.integers(0, n, e), rng.integers(0, n, e)], axis=1)
    return age, stage0, edges


def _gpu_episode(age, stage0, edges, seed, mode, r2, m, weights=None, **engine_kw):
    from agenttorch_b200.engine import NoiseSpec, SeirmEngine, SeirmSpec, SeirmState
    from agenttorch_b200.graph import build_contact_graph, mean_interactions_from_age

    dev = torch.device("cuda:0")
    n = len(age)
    ei = torch.from_numpy(edges.astype(np.int64)).t().contiguous()
    attr = None
    if weights is not None:
        attr = torch.stack((torch.ones(len(edges)), torch.from_numpy(weights.astype(np.float32))))
    graph = build_contact_graph(ei, n, edge_attr=attr, mean_interactions=mean_interactions_from_age(torch.from_numpy(age))).to(dev)
    eng = SeirmEngine(graph, SeirmSpec(num_steps=T), **engine_kw)
    st = SeirmState.from_initial(stage0.astype(np.float32), age, T, dev)
    r2_t = torch.tensor(r2, dtype=torch.float32, requires_grad=True)
    eye = torch.eye(3)
    r_steps = torch.stack([(r2_t * eye[int(t / 7)]).sum() for t in range(T)])
    m_t = torch.tensor(m, dtype=torch.float32, device=dev)
    sv = eng.run_episode(st, r_steps.to(dev), m_t, T, NoiseSpec(seed=seed, isolation=mode), want_grad=True)
    gI = torch.linspace(0.5, 1.5, T, device=dev)
    gD = torch.linspace(2.0, 1.0, T, device=dev)
    gr, gm = eng.episode_backward(sv, T, gI, gD, m_t)
    eng.check_errors()
    r_steps.backward(gr.cpu())
    return st, r2_t.grad.numpy(), gm.cpu().numpy(), eng


def _oracle(age, stage0, edges, seed, mode, r2, m, weights=None):
    U, W = make_noise(seed, mode, len(age), T)
    port = SeirmPort(age, stage0, edges, edge_weight=weights)
    gI = np.linspace(0.5, 1.5, T, dtype=np.float32)
    gD = np.linspace(2.0, 1.0, T, dtype=np.float32)
    return episode_with_grads(port, W, U, gI, gD, r2=r2, m=m)


def _compare(st, gr2, gm, ref):
    assert np.array_equal(st.stage.cpu().numpy().astype(np.float64), ref["stage"])                 # bit-exact
    assert np.array_equal(st.next_stage_time.cpu().numpy().astype(np.float32), ref["next_stage_time"])
    assert np.array_equal(st.infected_time.cpu().numpy().astype(np.float32), ref["infected_time"])
    assert rel_err(ref["daily_infected"], st.daily_infected.cpu().numpy()) <= 1e-5              # float state: 1e-5 relative
    assert rel_err(ref["daily_deaths"], st.daily_deaths.cpu().numpy()) <= 1e-5
    assert rel_err(ref["grad_R2"], gr2) <= 1e-4                                                   # gradients: 1e-4 relative
    assert rel_err(ref["grad_M"], gm) <= 1e-4


R2, M = np.array([[0.30], [0.25], [0.20]], dtype=np.float32), np.array([0.2], dtype=np.float32)


@pytest.mark.parametrize("mode", ["hard", "soft"])
@pytest.mark.parametrize("kw", [dict(), dict(block_size=2048, tile_rows=512), dict(layout="bcsr"), dict(layout="bcsr", block_size=2048),
                                dict(hash_words=128), dict(edges_per_chunk=512)])
def test_tasks_larger_than_one_chunk(mode, kw):
    """average in-degree 45: a 32-row task holds ~1440 edges = 2-3 TMA chunks of 512/768 edges"""
    age, stage0, edges = _inputs(6000, 45.0, 3)
    ref = _oracle(age, stage0, edges, 77, mode, R2 * 0.3, M)
    st, gr2, gm, eng = _gpu_episode(age, stage0, edges, 77, mode, R2 * 0.3, M, **kw)
    assert float(st.daily_infected.sum()) > 50
    _compare(st, gr2, gm, ref)


@pytest.mark.parametrize("kw", [dict(), dict(block_size=1024), dict(hash_words=64)])
def test_weighted_edges(kw):
    age, stage0, edges = _inputs(4000, 7.0, 4)
    w = np.random.default_rng(9).uniform(0.25, 2.0, len(edges)).astype(np.float32)
    ref = _oracle(age, stage0, edges, 5, "hard", R2 * 4, M, weights=w)
    st, gr2, gm, _ = _gpu_episode(age, stage0, edges, 5, "hard", R2 * 4, M, weights=w, **kw)
    _compare(st, gr2, gm, ref)


def test_100k_agents_against_the_oracle():
    age, stage0, edges = _inputs(100_000, 20.0, 11, frac_e=0.01, frac_i=0.0)
    r2 = np.array([[4.75], [4.75], [4.75]], dtype=np.float32)
    for mode in ("hard", "soft"):
        ref = _oracle(age, stage0, edges, 21, mode, r2, np.array([0.12], dtype=np.float32))
        st, gr2, gm, _ = _gpu_episode(age, stage0, edges, 21, mode, r2, np.array([0.12], dtype=np.float32))
        _compare(st, gr2, gm, ref)


def test_layouts_grids_and_reruns_are_bitwise_identical():
    age, stage0, edges = _inputs(50_000, 12.0, 8)
    base = None
    for kw in (dict(), dict(), dict(grid=5), dict(warps_per_cta=4), dict(block_size=8192), dict(warps_per_cta=24, tile_rows=256),
               dict(bank_aware=False), dict(layout="bcsr"), dict(layout="bcsr", grid=5), dict(layout="bcsr", warps_per_cta=4),
               dict(layout="bcsr", block_size=8192), dict(hash_words=256), dict(layout="bcsr", tasks_per_group=2)):
        st, gr2, gm, _ = _gpu_episode(age, stage0, edges, 3, "soft", R2 * 10, M, **kw)
        cur = (st.stage.cpu(), st.next_stage_time.cpu(), st.daily_infected.cpu(), st.daily_deaths.cpu())
        if base is None:
            base = cur
        else:
            for a, b in zip(base, cur):
                assert torch.equal(a, b), kw
    assert float(base[2].sum()) > 100


def test_edge_cases_empty_graph_and_no_infection():
    from agenttorch_b200.engine import NoiseSpec, SeirmEngine, SeirmSpec, SeirmState
    from agenttorch_b200.graph import build_contact_graph

    dev = torch.device("cuda:0")
    n = 100
    g = build_contact_graph(torch.zeros((2, 0), dtype=torch.long), n).to(dev)
    eng = SeirmEngine(g, SeirmSpec(num_steps=T))
    stage0 = np.zeros(n, dtype=np.float32)
    stage0[:5] = 1.0
    stage0[5:8] = 2.0
    st = SeirmState.from_initial(stage0, np.zeros(n), T, dev)
    eng.run_episode(st, torch.full((T,), 14.25, device=dev), torch.tensor([0.5], device=dev), T, NoiseSpec(seed=1), want_grad=True)
    assert float(st.daily_infected.sum()) == 0.0                       # nobody can be exposed without edges
    assert st.stage.cpu().tolist()[:8] == [3.0] * 8                    # initial E and I agents still progress to R
    dd = st.daily_deaths.cpu().numpy()
    assert dd[5] == pytest.approx(1.5) and dd[8] == pytest.approx(2.5) and dd.sum() == pytest.approx(4.0)


@pytest.mark.parametrize("n", [1_000_000, 8_400_000])       # BASELINE configs[1] and configs[2] sizes (d = 20)
def test_invariants_at_baseline_sizes(n):
    """size-independent properties where the oracle is too slow: conservation, monotone stages, series = stage counts"""
    from agenttorch_b200.engine import NoiseSpec, SeirmEngine, SeirmSpec, SeirmState
    from agenttorch_b200.graph import mean_interactions_from_age, synthetic_contact_graph, synthetic_population

    dev = torch.device("cuda:0")
    age, stage0 = synthetic_population(n, 1, device=dev)
    graph = synthetic_contact_graph(n, 20.0, 2, device=dev, mean_interactions=mean_interactions_from_age(age))
    eng = SeirmEngine(graph, SeirmSpec(num_steps=T))
    st = SeirmState.from_initial(stage0, age, T, dev)
    s0 = st.stage.clone()
    sv = eng.run_episode(st, torch.full((T,), 14.25, device=dev), torch.tensor([0.12], device=dev), T,
                         NoiseSpec(seed=4, isolation="hard"), want_grad=True, cuda_graph=True)
    stage = st.stage
    assert bool(((stage == 0) | (stage == 1) | (stage == 2) | (stage == 3)).all())
    assert bool((stage >= s0).all())
    newly = int(((s0 == 0) & (stage > 0)).sum())
    assert float(st.daily_infected.sum()) == float(newly)               # every exposure is counted exactly once
    deaths = float(sv["death_count"].sum())
    assert deaths == float(((stage == 3)).sum()) and float(st.daily_deaths.sum()) == pytest.approx(0.12 * deaths, rel=1e-5)
    exposed_at = st.infected_time[(s0 == 0) & (stage > 0)].long()
    hist = torch.bincount(exposed_at, minlength=T)[:T].float()
    assert torch.equal(hist, st.daily_infected)                         # infected_time records the exposure step
    eng.check_errors()


def test_fused_episode_with_intervention_substeps(tmp_path):
    """A covid config WITH the quarantine / testing substeps and the default (in-kernel) isolation draw runs as one fused
    episode: SEIRM step launches + one in-place launch per intervention block after every step.  Reference fixture
    `interventions_1500_softopen`; the same comparison runs on the emulated libraries in tests/test_seirm_emulated.py.
    (Kept last in the last GPU file: it was written after the round's GPU budget was spent.)"""
    from agenttorch_b200.core.dataloader import DataLoader, LoadPopulation
    from agenttorch_b200.core.runner import Runner
    from agenttorch_b200.models import covid
    from agenttorch_b200.models.covid.interventions import add_intervention_substeps
    from agenttorch_b200.populations import write_population_folder
    from tests.helpers import load_interventions_golden

    g = load_interventions_golden("interventions_1500_softopen")
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
    dev = runner.initializer.device
    assert dev.type == "cuda"
    tf, pf = runner.initializer.transition_function, runner.initializer.policy_function
    for mod in (tf["0"]["new_transmission"], pf["2"]["citizens"]["start_compliance"], pf["2"]["citizens"]["break_compliance"],
                pf["3"]["citizens"]["accept_test"]):
        mod.advance_seed = False
    r2 = torch.tensor(g["r2"], dtype=torch.float32, device=dev, requires_grad=True)
    m = torch.tensor(g["m"], dtype=torch.float32, device=dev, requires_grad=True)
    runner._set_parameters({"initializer.transition_function.0.new_transmission.learnable_args.R2": r2,
                            "initializer.transition_function.1.seirm_progression.learnable_args.M": m})
    Tn = g["num_steps"]
    cit = runner.state["agents"]["citizens"]
    env = runner.state["environment"]
    snapshot = {"agents": {"citizens": {k: v.detach().clone() for k, v in cit.items() if torch.is_tensor(v) and v.is_floating_point()}},
                "environment": {k: env[k].detach().clone() for k in ("daily_infected", "daily_deaths")}}
    from agenttorch_b200.models.covid.context import get_context

    # twice: the first call captures the SEIRM launches AND the intervention launches of the episode into one CUDA graph (the
    # hook's noise seeds live in device words), the second restores the state and replays it
    for rep in range(2):
        if rep:
            runner.restore_state(snapshot)
            r2 = torch.tensor(g["r2"], dtype=torch.float32, device=dev, requires_grad=True)
            m = torch.tensor(g["m"], dtype=torch.float32, device=dev, requires_grad=True)
            runner._set_parameters({"initializer.transition_function.0.new_transmission.learnable_args.R2": r2,
                                    "initializer.transition_function.1.seirm_progression.learnable_args.M": m})
        runner.step(Tn)
        assert runner.perf_stats["fused_episodes"] == rep + 1
        cit = runner.state["agents"]["citizens"]
        for k in ("disease_stage", "next_stage_time", "is_quarantined", "quarantine_start_date", "awaiting_test_result",
                  "test_result_date", "test_re_eligble_date"):
            got = cit[k].detach().cpu().numpy().reshape(-1).astype(np.float64)
            assert np.array_equal(got, g["ref_" + k + "_history"][Tn - 1].astype(np.float64)), (rep, k)
        env = runner.state["environment"]
        di, dd = env["daily_infected"], env["daily_deaths"]
        assert rel_err(g["ref_daily_infected"], di.detach().cpu().numpy()) <= 1e-5
        assert rel_err(g["ref_daily_deaths"], dd.detach().cpu().numpy()) <= 1e-5
        loss = (torch.from_numpy(g["gI"]).to(dev) * di.reshape(-1)).sum() + (torch.from_numpy(g["gD"]).to(dev) * dd.reshape(-1)).sum()
        loss.backward()
        assert rel_err(g["ref_grad_R2"], r2.grad.cpu().numpy()) <= 1e-4 and rel_err(g["ref_grad_M"], m.grad.cpu().numpy()) <= 1e-4
    engine = get_context(runner.config, runner.state, tf["0"]["new_transmission"].input_variables).engine
    assert len(engine._graphs) == 1                                  # captured once (with the hook's launches), replayed once
