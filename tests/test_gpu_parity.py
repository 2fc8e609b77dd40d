"""GPU parity: the CUDA step kernel (through the C ABI) vs the real reference's golden outputs and
vs the CPU oracle on seeded synthetic inputs.  Integer state bit-exact; float series 1e-5; grads 1e-4."""
import numpy as np
import pytest
import torch

from tests.helpers import GOLDEN_CASES, TABLE_CASES, load_golden, make_noise, rel_err

pytestmark = pytest.mark.gpu


def _r_steps_and_grad(r2_np, T, num_weeks=3):
    """per-step R exactly as the reference computes it (transition.py:136-143), with autograd"""
    r2 = torch.tensor(r2_np, dtype=torch.float32, requires_grad=True)
    eye = torch.eye(num_weeks)
    r_steps = torch.stack([(r2 * eye[int(t / 7)]).sum() for t in range(T)])
    return r2, r_steps


def _run_engine(g, *, packed, block_size=None, use_tensor_noise=False, cuda_graph=False, grid=0, tasks_per_group=0, hashed=None, hash_words=None,
                layout="auto", tile_rows=None, warps_per_cta=0):
    from agenttorch_b200.engine import NoiseSpec, SeirmEngine, SeirmSpec, SeirmState
    from agenttorch_b200.graph import build_contact_graph, mean_interactions_from_age

    dev = torch.device("cuda:0")
    T = g["num_steps"]
    n = len(g["age"])
    age = torch.from_numpy(g["age"].astype(np.int64))
    mi = mean_interactions_from_age(age)
    edges = torch.from_numpy(g["edges"].astype(np.int64)).t().contiguous()
    graph = build_contact_graph(edges, n, mean_interactions=mi).to(dev)
    spec = SeirmSpec(num_steps=T)
    eng = SeirmEngine(graph, spec, block_size=block_size, grid=grid, tasks_per_group=tasks_per_group, hashed=hashed, hash_words=hash_words,
                      layout=layout, tile_rows=tile_rows, warps_per_cta=warps_per_cta)
    st = SeirmState.from_initial(g["stage0"].astype(np.float32), g["age"], T, dev, packed=packed)
    if use_tensor_noise:
        U, W = make_noise(g["seed"], g["mode"], n, T, float(g["p_isolate"]))
        noise = NoiseSpec(seed=0, isolation="tensor", uniform="tensor",
                          isolation_tensor=torch.from_numpy(W).to(dev), uniform_tensor=torch.from_numpy(U).to(dev))
    else:
        noise = NoiseSpec(seed=g["seed"], isolation=g["mode"], p_isolate=float(g["p_isolate"]))
    r2, r_steps = _r_steps_and_grad(g["r2"], T)
    m = torch.tensor(g["m"], dtype=torch.float32)
    sv = eng.run_episode(st, r_steps.to(dev), m.to(dev), T, noise, want_grad=True, cuda_graph=cuda_graph)
    gr, gm = eng.episode_backward(sv, T, torch.from_numpy(g["gI"]).to(dev), torch.from_numpy(g["gD"]).to(dev), m.to(dev))
    eng.check_errors()
    r_steps.backward(gr.cpu())
    return st, r2.grad.numpy(), gm.cpu().numpy(), eng


def _check(g, st, grad_r2, grad_m):
    assert np.array_equal(st.stage.cpu().numpy().astype(np.float64), g["ref_stage"])
    assert np.array_equal(st.next_stage_time.cpu().numpy().astype(np.int64), g["ref_next_stage_time"].astype(np.int64))
    assert np.array_equal(st.infected_time.cpu().numpy().astype(np.int64), g["ref_infected_time"].astype(np.int64))
    assert rel_err(g["ref_daily_infected"], st.daily_infected.cpu().numpy()) <= 1e-5   # north star: float state 1e-5 rel
    assert rel_err(g["ref_daily_deaths"], st.daily_deaths.cpu().numpy()) <= 1e-5
    assert rel_err(g["ref_grad_R2"], grad_r2) <= 1e-4                                  # north star: gradients 1e-4 rel
    assert rel_err(g["ref_grad_M"], grad_m) <= 1e-4


@pytest.mark.parametrize("case", GOLDEN_CASES)
@pytest.mark.parametrize("packed,block_size", [(True, None), (False, 1024), (True, 2048)])
@pytest.mark.parametrize("layout", ["stream", "bcsr"])
def test_episode_matches_reference_golden(case, packed, block_size, layout):
    """stream: flat per-block streams + fixed-point hit records (gather + update kernels), one or several source blocks.
    bcsr: block_size=None one column block, fused kernel; 1024/2048: column-blocked gather + update kernels."""
    g = load_golden(case)
    st, gr2, gm, eng = _run_engine(g, packed=packed, block_size=block_size, layout=layout,
                                   tile_rows=(256 if block_size == 1024 else None) if layout == "stream" else None)
    assert eng.graph.layout == (1 if layout == "stream" else 0)
    if block_size is not None and len(g["age"]) > block_size:
        assert eng.graph.n_blocks > 1
    _check(g, st, gr2, gm)


@pytest.mark.parametrize("case", ["synth3k_soft", "astoria_hard", "astoria_soft"])
@pytest.mark.parametrize("kw", [dict(warps_per_cta=24), dict(warps_per_cta=16, grid=3), dict(tile_rows=64, block_size=4096, grid=301),
                                dict(tasks_per_group=2)])
def test_stream_layout_kernel_variants(case, kw):
    """the stream gather's register-budget variants (32 / 24 / 16 warps with 3 / 4 / 6 units in flight), grids that make CTAs
    span source blocks, tiny tiles"""
    g = load_golden(case)
    st, gr2, gm, eng = _run_engine(g, packed=True, layout="stream", **kw)
    assert eng.graph.layout == 1
    _check(g, st, gr2, gm)


@pytest.mark.parametrize("case", GOLDEN_CASES)
@pytest.mark.parametrize("hash_words", [64, 32])
def test_hashed_filter_mode(case, hash_words):
    """hashed shared-memory filter (used when the sources do not fit an exact bitmap): 1024 filter words here, so
    astoria's 37518 sources alias ~1.1 per bit and false positives must be rejected by the code lookup"""
    g = load_golden(case)
    # 2048 / 1024 filter bits: heavy aliasing; with 32 words astoria has > 24 sources per bit, which switches on the
    # exact-bitmap check in front of the code gather
    st, gr2, gm, eng = _run_engine(g, packed=True, hash_words=hash_words)
    assert eng.graph.hash_words == hash_words and eng.graph.n_blocks == 1
    _check(g, st, gr2, gm)


@pytest.mark.parametrize("case", ["synth3k_soft", "astoria_hard"])
def test_tensor_noise_and_cuda_graph_paths(case):
    g = load_golden(case)
    for layout in ("stream", "bcsr"):
        st, gr2, gm, _ = _run_engine(g, packed=True, use_tensor_noise=True, tasks_per_group=1, layout=layout)
        _check(g, st, gr2, gm)
        st, gr2, gm, eng = _run_engine(g, packed=True, block_size=1024, cuda_graph=True, grid=7, tasks_per_group=8, layout=layout)
        _check(g, st, gr2, gm)
        assert eng.launch_count > 0


@pytest.mark.parametrize("case", TABLE_CASES)
@pytest.mark.parametrize("kw", [dict(), dict(block_size=1024, tile_rows=256)])
def test_layers_and_table_gradients_match_reference_golden(case, kw):
    """multi-layer contact networks (per-edge network type -> per-edge I_bar, transition.py:71-73) and dL/dSFSusceptibility,
    dL/dSFInfector against the real reference (oracle/ref_harness/make_golden.py round2)"""
    from agenttorch_b200.engine import NoiseSpec, SeirmEngine, SeirmSpec, SeirmState
    from agenttorch_b200.graph import build_contact_graph, mean_interactions_from_age

    dev = torch.device("cuda:0")
    g = load_golden(case)
    T, n = g["num_steps"], len(g["age"])
    age = torch.from_numpy(g["age"].astype(np.int64))
    edges = torch.from_numpy(g["edges"].astype(np.int64)).t().contiguous()
    attr = None
    if "edge_type" in g:
        attr = torch.stack((torch.from_numpy(g["edge_type"]).float(), torch.ones(edges.shape[1])))
    graph = build_contact_graph(edges, n, edge_attr=attr, mean_interactions=mean_interactions_from_age(age)).to(dev)
    assert (len(graph.i_bar_layers) == 3) == ("edge_type" in g)
    eng = SeirmEngine(graph, SeirmSpec(num_steps=T), **kw)
    eng.want_table_grad = True
    st = SeirmState.from_initial(g["stage0"].astype(np.float32), g["age"], T, dev)
    noise = NoiseSpec(seed=g["seed"], isolation=g["mode"], p_isolate=float(g["p_isolate"]))
    r2, r_steps = _r_steps_and_grad(g["r2"], T)
    m = torch.tensor(g["m"], dtype=torch.float32, device=dev)
    for cuda_graph in (False, True):
        fresh = SeirmState.from_initial(g["stage0"].astype(np.float32), g["age"], T, dev)
        for k in ("stage", "next_stage_time", "infected_time", "daily_infected", "daily_deaths"):
            getattr(st, k).copy_(getattr(fresh, k))
        sv = eng.run_episode(st, r_steps.detach().to(dev), m, T, noise, want_grad=True, cuda_graph=cuda_graph)
    gI, gD = torch.from_numpy(g["gI"]).to(dev), torch.from_numpy(g["gD"]).to(dev)
    gr, gm = eng.episode_backward(sv, T, gI, gD, m)
    eng.check_errors()
    r_steps.backward(gr.cpu())
    _check(g, st, r2.grad.numpy(), gm.cpu().numpy())
    g_sfs, g_sfi = eng.table_gradients(sv, T, gI, gD, m, r_steps.detach().to(dev), st.age)
    assert rel_err(g["ref_grad_SFS"], g_sfs.cpu().numpy()) <= 1e-4
    assert rel_err(g["ref_grad_SFI"][1:3], g_sfi.cpu().numpy()[1:3]) <= 1e-4      # [0]: fractional-stage sources, not tracked
    if g["mode"] != "soft":
        assert rel_err(g["ref_grad_SFI"], g_sfi.cpu().numpy()) <= 1e-4
