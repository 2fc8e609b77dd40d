"""GPU parity of the predator-prey kernels (libat_predprey.so) against

  * the golden trajectories the REAL reference produced (tests/golden/predprey_*.npz),
  * the CPU restatement (oracle/predprey_port.py) on larger seeded grids,
  * the known-answer vectors of the reference's tests/test_vmap_substeps.py, through this package's substep classes,
and size-independent properties at the BASELINE size (10 M agents on a 4096 x 4096 grid).
All integer / float32 / float64 state must match bit for bit."""
import numpy as np
import pytest
import torch

from tests.helpers import PREDPREY_CASES, PREDPREY_KEYS, load_predprey_golden

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _spec_from_golden(g):
    from agenttorch_b200.predprey import PredPreySpec

    pw, yw, pn, gn, rt = (float(v) for v in g["consts"])
    ap, ay = PredPreySpec.work_flags(g["update_outputs"])
    mx, my = int(g["bounds"][0]), int(g["bounds"][1])
    return PredPreySpec(mx, my, g["n_cells"], mx, pw, yw, pn, gn, rt, ap, ay)


def _state_from(spec, src, prefix="init_"):
    from agenttorch_b200.predprey import PredPreyState

    t = lambda k: torch.from_numpy(np.ascontiguousarray(src[prefix + k])).to(DEV)      # noqa: E731
    return PredPreyState.from_reference(spec, t("pred_pos"), t("pred_energy"), t("prey_pos"), t("prey_energy"),
                                        t("growth_stage"), t("growth_countdown"))


def _assert_equal(st, spec, want, where):
    got = st.to_reference(spec)
    for k in PREDPREY_KEYS:
        a, b = got[k].cpu().numpy(), np.asarray(want[k])
        assert a.dtype == b.dtype, (where, k, a.dtype, b.dtype)
        assert np.array_equal(a, b), (where, k, int((a != b).sum()))


@pytest.mark.parametrize("case", PREDPREY_CASES)
@pytest.mark.parametrize("mode", ["substeps", "substeps_csr", "fused_graph", "fused_chunks", "fused_sorted"])
def test_kernels_match_reference_trajectory(case, mode):
    from agenttorch_b200.predprey import CsrAdjacency, PredPreyEngine

    import dataclasses

    g = load_predprey_golden(case)
    spec = _spec_from_golden(g)
    if mode == "fused_sorted":
        spec = dataclasses.replace(spec, tiled=True)         # grass stored in 8 x 16 tiles, agents sorted by tile address
    adj = CsrAdjacency(torch.from_numpy(g["adjacency"].astype(np.int64)), DEV) if mode == "substeps_csr" else None
    eng = PredPreyEngine(spec, DEV, adjacency=adj)
    st = _state_from(spec, g)
    T, seed = g["steps"], g["seed"]
    if mode.startswith("substeps"):
        for t in range(T):
            eng.move(st, t, seed); eng.eat(st); eng.hunt(st); eng.grow(st)
            _assert_equal(st, spec, {k: g["traj_" + k][t] for k in PREDPREY_KEYS}, (case, mode, t))
    elif mode == "fused_sorted":
        st.sort_by_cell(spec)                            # slots ordered by cell; the draws still follow the agent numbers
        addr = spec.storage_index(DEV)[st.prey_cell.long()]
        assert bool((addr[1:] >= addr[:-1]).all())
        eng.run(st, T, seed, cuda_graph=True, resort_every=5)
        _assert_equal(st, spec, {k: g["traj_" + k][T - 1] for k in PREDPREY_KEYS}, (case, mode))
        eng.run(st, 0, seed)
    elif mode == "fused_graph":
        eng.run(st, T, seed, cuda_graph=True)
        _assert_equal(st, spec, {k: g["traj_" + k][T - 1] for k in PREDPREY_KEYS}, (case, mode))
        st2 = _state_from(spec, g)                       # replay of the captured graph on another state object
        st.copy_(st2)
        eng.run(st, T, seed, cuda_graph=True)
        _assert_equal(st, spec, {k: g["traj_" + k][T - 1] for k in PREDPREY_KEYS}, (case, mode, "replay"))
    else:
        done = 0
        for chunk in (1, 3, 2, 5, T):                    # runner.step(k) called repeatedly: the draw clock continues
            chunk = min(chunk, T - done)
            if chunk <= 0:
                break
            eng.run(st, chunk, seed, t0=done, cuda_graph=(chunk % 2 == 1))
            done += chunk
            _assert_equal(st, spec, {k: g["traj_" + k][done - 1] for k in PREDPREY_KEYS}, (case, mode, done))


@pytest.mark.parametrize("shape,n_pred,n_prey,f64", [((300, 500), 40_000, 160_000, True), ((257, 129), 30_001, 50_003, False),
                                                     ((1024, 1024), 250_000, 750_000, False)])
def test_fused_steps_match_oracle_on_seeded_grids(shape, n_pred, n_prey, f64):
    from agenttorch_b200.predprey import PredPreyEngine, PredPreySpec
    from oracle.predprey_port import PredPreyConstants, PredPreyPort

    mx, my = shape
    rng = np.random.default_rng(mx * 7 + my)
    C = mx * my
    init = {
        "pred_pos": np.stack([rng.integers(0, mx, n_pred), rng.integers(0, my, n_pred)], 1),
        "prey_pos": np.stack([rng.integers(0, mx, n_prey), rng.integers(0, my, n_prey)], 1),
        "pred_energy": rng.uniform(-2, 30, (n_pred, 1)).astype(np.float32),
        "prey_energy": rng.uniform(-2, 40, (n_prey, 1)).astype(np.float32),
    }
    cd = np.where(rng.random(C) < 0.4, 0.0, rng.uniform(0.01, 12.0, C))
    init["growth_countdown"] = cd.reshape(-1, 1).astype(np.float64 if f64 else np.float32)
    init["growth_stage"] = (init["growth_countdown"] <= 0).astype(np.int64)
    consts = dict(pred_work=1.5, prey_work=2.25, prey_nutrition=6.5, grass_nutrition=3.3, regrowth_time=7.7)
    port = PredPreyPort(PredPreyConstants(mx, my, **consts), **init)
    spec = PredPreySpec(mx, my, C, mx, tiled=f64, **consts)
    eng = PredPreyEngine(spec, DEV)
    st = _state_from(spec, init, prefix="")
    T, seed = 12, 2026
    for t in range(T):
        port.step(t, seed)
    if f64:
        st.sort_by_cell(spec)
    eng.run(st, T, seed, resort_every=4 if f64 else 0)
    _assert_equal(st, spec, port.snapshot(), (shape, "fused"))
    alive_pred, alive_prey, grown = eng.census(st)
    assert alive_pred == int((port.pred_energy > 0).sum()) and alive_prey == int((port.prey_energy > 0).sum())
    assert grown == int(port.growth_stage.sum())
    assert 0 < alive_prey < n_prey and (port.prey_energy == 0).any()        # the run exercised deaths and hunts


# --------------------------------------------------- the reference's KATs through the substep classes
def _mk(cls, iv, ov):
    return cls(None, iv, ov, {"learnable": {}, "fixed": {}})


def _t(v, dtype=None):
    return torch.tensor(v, dtype=dtype, device=DEV)


def test_kat_move_substeps():                       # reference tests/test_vmap_substeps.py:40-226
    from agenttorch_b200.models.predator_prey.substeps.move import DecideMovement, FindNeighbors, UpdatePositions

    adj = torch.zeros(9, 9)
    for i in range(3):
        for j in range(3):
            k = 3 * i + j
            for di, dj in ((0, 1), (1, 0), (0, -1), (-1, 0)):
                if 0 <= i + di < 3 and 0 <= j + dj < 3:
                    adj[k, 3 * (i + di) + (j + dj)] = 1
    state = {
        "environment": {"bounds": torch.tensor([3, 3])},
        "network": {"agent_agent": {"predator_prey": {"adjacency_matrix": adj.to(DEV)}}},
        "agents": {
            "predator": {"coordinates": _t([[0, 0], [1, 1], [2, 2]]), "energy": _t([[10.0], [5.0], [0.0]]), "stride_work": _t([1.0])},
            "prey": {"coordinates": _t([[0, 1], [1, 0], [2, 1]]), "energy": _t([[8.0], [6.0], [4.0]]), "stride_work": _t([1.0])},
        },
    }
    find = _mk(FindNeighbors, {"bounds": "environment/bounds", "adj_grid": "network/agent_agent/predator_prey/adjacency_matrix",
                               "positions": "agents/predator/coordinates"}, ["possible_neighbors"])
    nb = find(state)["possible_neighbors"]
    assert len(nb) == 3 and len(nb[0]) == 2                                     # :135-147
    assert sorted(map(tuple, nb[0].tolist())) == [(0, 1), (1, 0)] and len(nb[2]) > 0
    decide = _mk(DecideMovement, {"positions": "agents/predator/coordinates", "energy": "agents/predator/energy"}, ["next_positions"])
    nxt = decide(state, {"possible_neighbors": nb})["next_positions"]
    assert nxt.shape == (3, 2)                                                  # :176-193
    assert not torch.equal(nxt[0], _t([0, 0])) and not torch.equal(nxt[1], _t([1, 1])) and torch.equal(nxt[2], _t([2, 2]))
    assert tuple(nxt[0].tolist()) in {(0, 1), (1, 0)}
    lists = {"possible_neighbors": [torch.tensor([[0, 1], [1, 0]]), torch.tensor([[0, 1], [1, 0], [1, 2], [2, 1]]), torch.tensor([[1, 2], [2, 1]])]}
    nxt2 = _mk(DecideMovement, decide.input_variables, ["next_positions"])(state, lists)["next_positions"]
    assert torch.equal(nxt2, nxt)                      # explicit lists (as the reference test passes them): same draws
    upd = _mk(UpdatePositions, {"prey_pos": "agents/prey/coordinates", "prey_energy": "agents/prey/energy",
                                "pred_pos": "agents/predator/coordinates", "pred_energy": "agents/predator/energy",
                                "prey_work": "agents/prey/stride_work", "pred_work": "agents/predator/stride_work"},
              ["prey_pos", "prey_energy", "pred_pos", "pred_energy"])
    action = {"prey": {"next_positions": _t([[0, 2], [1, 1], [2, 0]])}, "predator": {"next_positions": _t([[0, 1], [1, 2], [2, 2]])}}
    res = upd(state, action)                                                    # :196-226
    assert torch.equal(res["prey_pos"], action["prey"]["next_positions"]) and torch.equal(res["pred_pos"], action["predator"]["next_positions"])
    assert torch.equal(res["prey_energy"], _t([[7.0], [5.0], [3.0]])) and torch.equal(res["pred_energy"], _t([[9.0], [4.0], [-1.0]]))


def test_kat_eat_substeps():                        # reference tests/test_vmap_substeps.py:229-345
    from agenttorch_b200.models.predator_prey.substeps.eat import EatGrass, FindEatableGrass

    state = {
        "environment": {"bounds": torch.tensor([3, 3])},
        "agents": {"prey": {"coordinates": _t([[0, 0], [1, 1], [2, 2]]), "energy": _t([[10.0], [5.0], [2.0]])}},
        "objects": {"grass": {"growth_stage": _t([[1], [0], [1], [0], [1], [0], [1], [0], [1]]),
                              "growth_countdown": _t([[0], [5], [0], [10], [0], [15], [0], [8], [0]]),
                              "regrowth_time": _t([20.0]), "nutritional_value": _t([5.0])}},
    }
    find = _mk(FindEatableGrass, {"bounds": "environment/bounds", "positions": "agents/prey/coordinates",
                                  "grass_growth": "objects/grass/growth_stage"}, ["eatable_grass_positions"])
    eatable = find(state, {})["eatable_grass_positions"]
    assert len(eatable) == 3                          # the three prey stand on grown grass (the reference asserts 5, :292, which its own code cannot return)
    for pos in eatable:
        assert state["objects"]["grass"]["growth_stage"][int(pos[1]) + 3 * int(pos[0])].item() == 1
    eat = _mk(EatGrass, {"bounds": "environment/bounds", "prey_pos": "agents/prey/coordinates", "energy": "agents/prey/energy",
                         "nutrition": "objects/grass/nutritional_value", "grass_growth": "objects/grass/growth_stage",
                         "growth_countdown": "objects/grass/growth_countdown", "regrowth_time": "objects/grass/regrowth_time"},
              ["energy", "grass_growth", "growth_countdown"])
    for action in ({"prey": {"eatable_grass_positions": [torch.tensor([0, 0]), torch.tensor([1, 1]), torch.tensor([2, 2])]}},
                   {"prey": {"eatable_grass_positions": eatable}}):
        res = eat(state, action)                                                # :304-345
        assert torch.equal(res["energy"], _t([[15.0], [10.0], [7.0]]))
        assert res["grass_growth"].reshape(-1).tolist() == [0, 0, 1, 0, 0, 0, 1, 0, 0]
        assert res["growth_countdown"].reshape(-1).tolist() == [20, 5, 0, 10, 20, 15, 0, 8, 20]
    assert eat(state, {"prey": {"eatable_grass_positions": []}}) == {}


def test_kat_hunt_substeps():                       # reference tests/test_vmap_substeps.py:348-437
    from agenttorch_b200.models.predator_prey.substeps.hunt import FindTargets, HuntPrey

    state = {"agents": {
        "predator": {"coordinates": _t([[0, 0], [1, 1], [2, 2]]), "energy": _t([[10.0], [5.0], [2.0]])},
        "prey": {"coordinates": _t([[0, 0], [1, 1], [0, 1], [2, 1]]), "energy": _t([[8.0], [6.0], [4.0], [3.0]]),
                 "nutritional_value": _t([15.0])}}}
    find = _mk(FindTargets, {"prey_pos": "agents/prey/coordinates", "pred_pos": "agents/predator/coordinates"}, ["target_positions"])
    targets = find(state, {})["target_positions"]
    assert len(targets) == 2 and sorted(map(tuple, torch.stack(list(targets)).tolist())) == [(0, 0), (1, 1)]
    hunt = _mk(HuntPrey, {"prey_pos": "agents/prey/coordinates", "prey_energy": "agents/prey/energy", "pred_pos": "agents/predator/coordinates",
                          "pred_energy": "agents/predator/energy", "nutritional_value": "agents/prey/nutritional_value"},
               ["prey_energy", "pred_energy"])
    for action in ({"predator": {"target_positions": [torch.tensor([0, 0]), torch.tensor([1, 1])]}}, {"predator": {"target_positions": targets}}):
        res = hunt(state, action)
        assert res["prey_energy"].reshape(-1).tolist() == [0.0, 0.0, 4.0, 3.0]
        assert res["pred_energy"].reshape(-1).tolist() == [25.0, 20.0, 2.0]
    assert hunt(state, {"predator": {"target_positions": []}}) == {}


def test_kat_grow_substep():                        # reference tests/test_vmap_substeps.py:440-488
    from agenttorch_b200.models.predator_prey.substeps.grow import GrowGrass

    state = {"objects": {"grass": {"growth_stage": _t([[0], [0], [0], [0], [1], [0], [1], [0], [1]]),
                                   "growth_countdown": _t([[5], [3], [1], [0], [0], [10], [0], [7], [0]])}}}
    grow = _mk(GrowGrass, {"grass_growth": "objects/grass/growth_stage", "growth_countdown": "objects/grass/growth_countdown"},
               ["grass_growth", "growth_countdown"])
    res = grow(state, {})
    assert torch.equal(res["growth_countdown"], _t([[4], [2], [0], [-1], [-1], [9], [-1], [6], [-1]]))
    assert torch.equal(res["grass_growth"], _t([[0], [0], [1], [1], [1], [0], [1], [0], [1]]))


# ------------------------------------------------------------------------- Runner: drop-in behaviour
def _runner(fused, **kw):
    from agenttorch_b200.models import predator_prey as pp
    from agenttorch_b200.models.predator_prey.synthetic import synthetic_config

    cfg = synthetic_config(device="cuda", **kw)
    cfg["simulation_metadata"]["cuda_params"]["fused_episode"] = fused
    r = pp.get_runner(cfg, pp.registry)
    r.init()
    return r


def _public(r):
    s = r.state
    return {"pred_pos": s["agents"]["predator"]["coordinates"], "pred_energy": s["agents"]["predator"]["energy"],
            "prey_pos": s["agents"]["prey"]["coordinates"], "prey_energy": s["agents"]["prey"]["energy"],
            "growth_stage": s["objects"]["grass"]["growth_stage"], "growth_countdown": s["objects"]["grass"]["growth_countdown"]}


def test_runner_generic_and_fused_paths_agree_with_the_oracle():
    from oracle.predprey_port import PredPreyConstants, PredPreyPort

    kw = dict(max_x=40, max_y=56, num_predators=300, num_prey=900, seed=5, num_steps=6)
    generic, fused = _runner(False, **kw), _runner(True, **kw)
    init = {k: v.cpu().numpy().copy() for k, v in _public(generic).items()}
    for k, v in _public(fused).items():
        assert np.array_equal(v.cpu().numpy(), init[k])                         # seeded population: same on both
    port = PredPreyPort(PredPreyConstants(40, 56), **init)
    generic.step(6)
    fused.step(2); fused.step(1); fused.step(3)                                # step(k) calls continue the same episode
    for t in range(6):
        port.step(t, 5)
    want = port.snapshot()
    for name, r in (("generic", generic), ("fused", fused)):
        for k, v in _public(r).items():
            assert np.array_equal(v.cpu().numpy().reshape(want[k].shape), want[k]), (name, k)
    assert fused.perf_stats["fused_episodes"] == 3 and generic.perf_stats["fused_episodes"] == 0
    assert len(fused.state_trajectory) == len(generic.state_trajectory) == 7
    # a user edit of the public state between calls is picked up by the fused path
    fused.state["agents"]["prey"]["coordinates"][:10] = 0
    generic.state["agents"]["prey"]["coordinates"] = generic.state["agents"]["prey"]["coordinates"].clone()
    generic.state["agents"]["prey"]["coordinates"][:10] = 0
    fused.step(2); generic.step(2)
    for k in PREDPREY_KEYS:
        assert torch.equal(_public(fused)[k].reshape(-1), _public(generic)[k].reshape(-1)), k
    fused.reset()
    assert np.array_equal(_public(fused)["prey_pos"].cpu().numpy(), init["prey_pos"])


def test_baseline_size_properties():
    """BASELINE configs[3]: 10 M agents, 4096 x 4096 cells — determinism, graph == eager, conservation-type checks."""
    from agenttorch_b200.predprey import PredPreyEngine, PredPreySpec, PredPreyState

    mx = my = 4096
    C, n_pred, n_prey = mx * my, 2_000_000, 8_000_000
    g = torch.Generator(device=DEV); g.manual_seed(1)
    spec = PredPreySpec(mx, my, C, mx)
    pos = lambda n: torch.stack([torch.randint(0, mx, (n,), generator=g, device=DEV), torch.randint(0, my, (n,), generator=g, device=DEV)], 1)   # noqa: E731
    cd = torch.where(torch.rand(C, generator=g, device=DEV) < 0.4, 0.0, torch.randint(1, 101, (C,), generator=g, device=DEV).float())
    st0 = PredPreyState.from_reference(spec, pos(n_pred), 30 + 71 * torch.rand(n_pred, 1, generator=g, device=DEV), pos(n_prey),
                                       40 + 61 * torch.rand(n_prey, 1, generator=g, device=DEV), (cd <= 0).long(), cd)
    eng = PredPreyEngine(spec, DEV)
    a, b, c = st0.clone(), st0.clone(), st0.clone()
    eng.run(a, 10, 99, cuda_graph=True)
    eng.run(b, 4, 99, cuda_graph=False); eng.run(b, 6, 99, t0=4, cuda_graph=True)
    for k in ("pred_cell", "pred_energy", "prey_cell", "prey_energy", "info", "countdown"):
        assert torch.equal(getattr(a, k), getattr(b, k)), k
    # tiled grass storage + slots sorted by tile address change nothing but the speed
    import dataclasses
    tspec = dataclasses.replace(spec, tiled=True)
    ref0 = st0.to_reference(spec)
    c = PredPreyState.from_reference(tspec, ref0["pred_pos"], ref0["pred_energy"], ref0["prey_pos"], ref0["prey_energy"],
                                     ref0["growth_stage"], ref0["growth_countdown"], sort=True)
    PredPreyEngine(tspec, DEV).run(c, 10, 99, cuda_graph=True, resort_every=4)
    for k in ("pred_cell", "pred_energy", "prey_cell", "prey_energy"):
        assert torch.equal(getattr(a, k), c.in_agent_order(k)), k
    c_info, c_cd = c.grass_in_cell_order(tspec)
    assert torch.equal(a.info, c_info) and torch.equal(a.countdown, c_cd)
    # grass: stage <=> countdown <= 0 after every grow; countdowns stay in [-(steps), regrowth]
    from agenttorch_b200.predprey import unpack_stage
    stage = unpack_stage(a.info, C).reshape(-1)
    assert int(a.info.max()) <= 1                                # no scratch bit survives a run
    assert torch.equal(stage == 1, a.countdown <= 0)
    assert float(a.countdown.max()) <= 100.0 and float(a.countdown.min()) >= -10.0
    # agents moved at most one lattice step per step (Manhattan distance <= 10), none left the grid
    for c0, c1 in ((st0.pred_cell, a.pred_cell), (st0.prey_cell, a.prey_cell)):
        dx, dy = (c1 // my - c0 // my).abs(), (c1 % my - c0 % my).abs()
        assert int((dx + dy).max()) <= 10 and int(c1.min()) >= 0 and int(c1.max()) < C
    # energies: predators only lose 1 per step and gain 20 per hunt; hunted prey are exactly 0
    d = a.pred_energy - st0.pred_energy
    assert float(d.min()) >= -10.001 and float(((d + 10) / 20).round().sub((d + 10) / 20).abs().max()) < 1e-3
    assert int((a.prey_energy == 0).sum()) > 0
    assert eng.census(a) == (int((a.pred_energy > 0).sum()), int((a.prey_energy > 0).sum()), int((stage == 1).sum()))


# ----------------------------------------------------------------------------------------- edge cases
@pytest.mark.parametrize("shape,n_pred,n_prey", [((7, 1), 5, 9), ((1, 9), 4, 4), ((6, 5), 0, 37), ((6, 5), 23, 0), ((2, 2), 50, 50),
                                                 ((33, 31), 1, 1)])
def test_edge_cases_match_oracle(shape, n_pred, n_prey):
    """Degenerate lattices (one row / one column: cells with 1 or 2 neighbours), a missing species, crowded cells, and
    cell counts that are not a multiple of 4 / 16 / 32 (ragged tail of the cell pass)."""
    from agenttorch_b200.predprey import PredPreyEngine, PredPreySpec
    from oracle.predprey_port import PredPreyConstants, PredPreyPort

    mx, my = shape
    C = mx * my
    rng = np.random.default_rng(C * 31 + n_pred)
    init = {
        "pred_pos": np.stack([rng.integers(0, mx, n_pred), rng.integers(0, my, n_pred)], 1).reshape(-1, 2),
        "prey_pos": np.stack([rng.integers(0, mx, n_prey), rng.integers(0, my, n_prey)], 1).reshape(-1, 2),
        "pred_energy": rng.uniform(-1, 6, (n_pred, 1)).astype(np.float32),
        "prey_energy": rng.uniform(-1, 9, (n_prey, 1)).astype(np.float32),
        "growth_countdown": np.where(rng.random(C) < 0.5, 0.0, rng.uniform(0.5, 3.0, C)).reshape(-1, 1),
    }
    init["growth_stage"] = (init["growth_countdown"] <= 0).astype(np.int64)
    consts = dict(pred_work=0.75, prey_work=1.25, prey_nutrition=2.5, grass_nutrition=1.5, regrowth_time=2.2)
    port = PredPreyPort(PredPreyConstants(mx, my, **consts), **init)
    spec = PredPreySpec(mx, my, C, mx, tiled=True, **consts)
    eng = PredPreyEngine(spec, DEV)
    st = _state_from(spec, init, prefix="")
    st.sort_by_cell(spec)
    for t in range(9):
        port.step(t, 77)
    eng.run(st, 4, 77, cuda_graph=False)
    eng.run(st, 5, 77, t0=4, cuda_graph=True)
    _assert_equal(st, spec, port.snapshot(), (shape, n_pred, n_prey))
    assert eng.census(st) == (int((port.pred_energy > 0).sum()), int((port.prey_energy > 0).sum()), int(port.growth_stage.sum()))


def test_invalid_inputs_fail_loudly():
    from agenttorch_b200.predprey import PredPreySpec, PredPreyState

    spec = PredPreySpec(4, 4, 16, 4)
    z = lambda *s, dt=torch.float32: torch.zeros(*s, dtype=dt, device=DEV)             # noqa: E731
    with pytest.raises(ValueError, match="outside the grid"):
        PredPreyState.from_reference(spec, torch.tensor([[4, 0]], device=DEV), z(1, 1), z(0, 2, dt=torch.int64), z(0, 1),
                                     z(16, 1, dt=torch.int64), z(16, 1))
    with pytest.raises(ValueError, match="one entry per cell"):
        PredPreyState.from_reference(spec, z(0, 2, dt=torch.int64), z(0, 1), z(0, 2, dt=torch.int64), z(0, 1),
                                     z(15, 1, dt=torch.int64), z(15, 1))
