"""Host-side mirror of the reference's plugin surface: Registry, config reader, Initializer, Controller, Runner API,
decorators, vmap.  Cases follow the reference's own tests where they exist (tests/test_initializer.py)."""
import os

import numpy as np
import pytest
import torch
import yaml

from agenttorch_b200.core import Registry, Runner, envs
from agenttorch_b200.core.controller import Controller
from agenttorch_b200.core.dataloader import DataLoader, LoadPopulation
from agenttorch_b200.core.helpers import copy_module, get_by_path, read_config, resolve_config, set_by_path, to_cpu
from agenttorch_b200.core.initializer import Initializer
from agenttorch_b200.core.substep import SubstepAction, SubstepTransition, SubstepTransitionMessagePassing, vmap
from agenttorch_b200.models import covid
from agenttorch_b200.models.covid.synthetic import synthetic_config
from agenttorch_b200.populations import write_population_folder


@pytest.fixture
def initializer():
    return Initializer({"simulation_metadata": {"device": "cpu"}}, object())


# ---- reference tests/test_initializer.py:28-143, same inputs and expectations
def test_initialize_from_default_broadcasting(initializer):
    r = initializer._initialize_from_default([0.0, 0.0], [1000, 2])
    assert r.shape == (1000, 2) and torch.all(r == 0)
    r = initializer._initialize_from_default([1.0, 2.0, 3.0], [500, 3])
    assert torch.allclose(r[0], torch.tensor([1.0, 2.0, 3.0])) and torch.all(r == r[0])
    r = initializer._initialize_from_default(5.0, [100, 2])
    assert r.shape == (100, 2) and torch.all(r == 5.0) and r.dtype == torch.float32
    r = initializer._initialize_from_default([1, 2, 3, 4], [100, 2])
    assert r.shape == (100, 2) and torch.all(r == 1.0)
    assert initializer._initialize_from_default("test_string", [10, 2]) == "test_string"
    assert torch.allclose(initializer._initialize_from_default([1.0, 2.0], None), torch.tensor([1.0, 2.0]))
    r = initializer._initialize_from_default([1.0, 2.0], [10, 5, 2])
    assert r.shape == (10, 5, 2) and torch.allclose(r[-1, -1], torch.tensor([1.0, 2.0]))
    assert initializer._initialize_from_default([], [10, 0]).shape == (10, 0)
    r = initializer._initialize_from_default([42.0], [5, 3])
    assert r.shape == (5, 3) and torch.all(r == 42.0)
    assert initializer._initialize_from_default([1.0], [2, 2]).device.type == "cpu"


def test_config_interpolation_and_resolvers(tmp_path):
    raw = {"simulation_metadata": {"num_agents": 10, "dir": "/pop", "file": "${simulation_metadata.dir}/age.pickle",
                                   "twice": "${multiply:${simulation_metadata.num_agents},2}", "half": "${divide:10,4}",
                                   "plus": "${sum:1,2}"},
           "state": {"shape": ["${simulation_metadata.num_agents}", 1], "nested": {"n": "${state.shape.0}"}}}
    cfg = resolve_config(raw)
    assert cfg["simulation_metadata"]["file"] == "/pop/age.pickle"
    assert cfg["simulation_metadata"]["twice"] == 20 and cfg["simulation_metadata"]["half"] == 2
    assert cfg["simulation_metadata"]["plus"] == 3
    assert cfg["state"]["shape"] == [10, 1] and cfg["state"]["nested"]["n"] == 10
    p = tmp_path / "c.yaml"
    p.write_text(yaml.safe_dump(raw))
    assert read_config(str(p)) == cfg
    with pytest.raises(ValueError):
        read_config(str(tmp_path / "c.json"))


def test_packaged_config_matches_the_reference_schema():
    here = os.path.dirname(covid.__file__)
    cfg = read_config(os.path.join(here, "yamls", "config.yaml"))
    meta = cfg["simulation_metadata"]
    assert meta["num_substeps_per_step"] == len(cfg["substeps"]) == 2
    tr = cfg["substeps"]["0"]["transition"]["new_transmission"]
    assert tr["arguments"]["R2"]["shape"] == [3, 1] and tr["arguments"]["R2"]["value"] == 4.75
    assert tr["output_variables"] == ["disease_stage", "next_stage_time", "infected_time", "daily_infected"]
    assert all(v in tr["input_variables"] for v in tr["output_variables"])      # controller.py:77-81 writes back by these keys
    assert cfg["substeps"]["1"]["transition"]["seirm_progression"]["arguments"]["M"]["value"] == 0.12
    ref_yaml = "/root/reference/agent_torch/models/covid/yamls/config.yaml"
    if os.path.exists(ref_yaml):      # build container only: the reference's own YAML must load through our reader
        ref = read_config(ref_yaml)
        assert ref["substeps"]["0"]["transition"]["new_transmission"]["input_variables"] == tr["input_variables"]
        assert ref["state"]["environment"]["SFInfector"]["value"] == cfg["state"]["environment"]["SFInfector"]["value"]


def test_registry_is_shared_and_wired_like_the_reference():
    reg = covid.registry
    assert set(reg.helpers) == {"transition", "observation", "policy", "initialization", "network"}
    assert {"new_transmission", "seirm_progression"} <= set(reg.transition_helpers)
    assert "make_isolation_decision" in reg.policy_helpers and "network_from_file" in reg.network_helpers
    for name in ("read_from_file", "get_lam_gamma_integrals", "get_mean_agent_interactions", "get_infected_time",
                 "get_next_stage_time", "load_population_attribute", "initialize_id"):
        assert name in reg.initialization_helpers

    @Registry.register_substep("my_policy", key="policy")
    class MyPolicy(SubstepAction):
        def forward(self, state, observation):
            return {}

    assert Registry().policy_helpers["my_policy"] is MyPolicy       # class-level tables
    del Registry.helpers["policy"]["my_policy"]
    with pytest.raises(KeyError):
        reg.register(object, "x", key="nope")


def test_substep_constructor_contract():
    cfg = {"simulation_metadata": {"calibration": True}}

    class T(SubstepTransition):
        def forward(self, state, action):
            return {}

    t = T(cfg, {"a": "agents/x/a"}, ["a"], {"learnable": {"R2": torch.full((3, 1), 4.75)}, "fixed": {"k": torch.ones(1)}})
    assert isinstance(t.learnable_args, torch.nn.ParameterDict) and t.calibrate_R2.requires_grad and t.calibrate_R2.is_leaf
    assert torch.equal(t.calibrate_R2.detach(), torch.full((3, 1), 4.75)) and set(t.args) == {"R2", "k"}
    t2 = T({"simulation_metadata": {"calibration": False}}, {}, [], {"learnable": {}, "fixed": {}})
    assert not hasattr(t2, "calibrate_R2")


def test_native_message_passing_matches_gather_scatter_add():
    class MP(SubstepTransitionMessagePassing):
        def message(self, x_i, x_j, w):
            return (x_i + 2 * x_j) * w.view(-1, 1)

        def forward(self, state, action):
            return {}

    mp = MP({"simulation_metadata": {"calibration": False}}, {}, [], {"learnable": {}, "fixed": {}})
    x = torch.arange(8.0).view(4, 2)
    ei = torch.tensor([[0, 1, 3, 3], [1, 1, 0, 2]])
    w = torch.tensor([1.0, 2.0, 3.0, 4.0])
    out = mp.propagate(ei, x=x, w=w)
    ref = torch.zeros(4, 2)
    for e in range(4):
        ref[ei[1, e]] += (x[ei[1, e]] + 2 * x[ei[0, e]]) * w[e]
    assert torch.allclose(out, ref)


def test_helpers_and_controller_paths():
    st = {"a": {"b": torch.ones(2)}, "net": {"adj": (torch.zeros(2, 3), torch.ones(2, 3))}, "k": 3, "current_substep": "0"}
    assert get_by_path(st, ["a", "b"]) is st["a"]["b"]
    set_by_path(st, ["a", "c"], torch.zeros(1))
    cp = copy_module(st)
    assert cp["a"]["b"] is not st["a"]["b"] and torch.equal(cp["a"]["b"], st["a"]["b"])
    assert cp["net"]["adj"][0] is st["net"]["adj"][0]             # edge lists are shared, not deep-copied (Q8)
    assert to_cpu(st)["k"] == 3

    cfg = {"simulation_metadata": {"num_substeps_per_step": 2},
           "substeps": {"0": {"transition": {"f": {"input_variables": {"v": "a/b"}}}, "policy": {"x": None}, "observation": {"x": None}}}}

    class F(torch.nn.Module):
        def forward(self, state, action):
            return {"v": state["a"]["b"] + 1}

    ctl = Controller(cfg)
    nxt = ctl.progress(st, {}, {"0": {"f": F()}})
    assert nxt["current_substep"] == "1" and st["current_substep"] == "0"
    assert torch.equal(nxt["a"]["b"], torch.full((2,), 2.0)) and torch.equal(st["a"]["b"], torch.ones(2))
    assert ctl.act(st, None, {}, "x") is None and ctl.observe(st, {}, "x") is None


def test_with_behavior_and_vmap():
    from agenttorch_b200.models.covid.substeps.new_transmission.action import MakeIsolationDecision

    cfg = {"simulation_metadata": {"calibration": False, "device": "cpu", "EXECUTION_MODE": "llm", "num_agents": 5}}
    pol = MakeIsolationDecision(cfg, {}, ["isolation_decision"], {"learnable": {}, "fixed": {}})
    out = pol(None, None)["isolation_decision"]
    assert out.shape == (5, 1) and pol.draws_uniform()
    pol.behavior = lambda obs: torch.full((5, 1), 0.25)
    assert not pol.draws_uniform() and torch.all(pol(None, None)["isolation_decision"] == 0.25)
    MakeIsolationDecision.set_behavior(lambda obs: torch.zeros(5, 1))
    try:
        assert MakeIsolationDecision(cfg, {}, ["d"], {"learnable": {}, "fixed": {}}).behavior is not None
    finally:
        MakeIsolationDecision._class_behavior = None

    @vmap(agent_args=["pos", "alive"], shared_args=["scale"], outputs=["pos"])
    class Move(SubstepTransition):
        def forward(self, state, action=None):
            return {"pos": state["pos"] * state["scale"] * state["alive"] + self.args_k(state)}

        def args_k(self, state):
            return state["k"]

    mv = Move({"simulation_metadata": {"calibration": False}}, {}, ["position"], {"learnable": {}, "fixed": {"k": torch.tensor(1.0)}})
    state = {"agents": {"a": {"pos": torch.arange(6.0).view(3, 2), "alive": torch.tensor([1.0, 0.0, 1.0])}},
             "environment": {"scale": torch.tensor(2.0)}}
    res = mv(state)
    assert set(res) == {"position"}
    assert torch.allclose(res["position"], torch.tensor([[1.0, 3.0], [1.0, 1.0], [9.0, 11.0]]))


def test_runner_api_on_cpu_builds_state_and_refuses_to_step(tmp_path):
    folder = write_population_folder(str(tmp_path / "pop"), [0, 1, 5, 3, 2, 4], [0, 1, 0, 2, 0, 0], [[0, 1], [1, 2], [3, 4], [1, 0]])
    pop = LoadPopulation(folder)
    assert pop.population_size == 6 and pop.age.dtype == torch.float32
    dl = DataLoader(covid, pop)
    cfg = dl.get_config()
    assert cfg["simulation_metadata"]["num_agents"] == 6 and cfg["simulation_metadata"]["population_dir"] == folder
    with open(os.path.join(os.path.dirname(covid.__file__), "yamls", "config.yaml")) as f:
        assert yaml.safe_load(f)["simulation_metadata"]["num_agents"] == 37518        # packaged YAML is NOT rewritten
    cfg["simulation_metadata"]["device"] = "cpu"
    runner = Runner(cfg, covid.registry)
    runner.init()
    cit = runner.state["agents"]["citizens"]
    assert cit["disease_stage"].reshape(-1).tolist() == [0, 1, 0, 2, 0, 0]
    assert cit["infected_time"].reshape(-1).tolist() == [500, -1, 500, -3, 500, 500]      # utils.py:30-41
    assert cit["next_stage_time"].reshape(-1).tolist() == [500, 3, 500, 5, 500, 500]      # utils.py:44-57
    assert runner.state["environment"]["mean_interactions"].reshape(-1).tolist() == [2, 3, 4, 3, 3, 3]
    assert runner.state["environment"]["lam_gamma_integrals"].shape == (21,)
    ei, ea = runner.state["network"]["agent_agent"]["infection_network"]["adjacency_matrix"]
    assert ei.tolist() == [[0, 1, 3, 1], [1, 2, 4, 0]] and ea.shape == (2, 4) and bool((ea == 1).all())
    names = dict(runner.named_parameters())
    assert "initializer.transition_function.0.new_transmission.learnable_args.R2" in names
    assert "initializer.transition_function.1.seirm_progression.learnable_args.M" in names
    r2 = torch.ones(3, requires_grad=True)
    runner._set_parameters({"initializer.transition_function.0.new_transmission.learnable_args.R2": r2})
    tr = runner.initializer.transition_function["0"]["new_transmission"]
    assert tr.calibrate_R2 is r2
    assert [float(tr.effective_r(t)) for t in (0, 7, 20)] == [1.0, 1.0, 1.0]
    assert tr.r_steps(21).shape == (21,)
    with pytest.raises(AssertionError):
        runner._set_parameters({"initializer.transition_function.0.new_transmission.learnable_args.R2": torch.ones(3)})
    with pytest.raises(RuntimeError, match="CUDA"):          # no CPU fallback for the kernels
        runner.step(1)
    assert len(runner.state_trajectory) >= 1 and "environment" in runner.state_trajectory[0][0]


def test_reference_r_quirk_q2():
    cfg = synthetic_config(256, 3.0, device="cpu")
    runner = Runner(cfg, covid.registry)
    runner.init()
    tr = runner.initializer.transition_function["0"]["new_transmission"]
    # YAML R2 is (NUM_WEEKS,1): the reference's broadcast makes R the SUM of all weeks at every step
    assert float(tr.effective_r(0)) == pytest.approx(14.25) and float(tr.effective_r(15)) == pytest.approx(14.25)
    tr.calibrate_R2 = torch.tensor([1.0, 2.0, 3.0], requires_grad=True)
    assert [float(tr.effective_r(t)) for t in (0, 7, 14)] == [1.0, 2.0, 3.0]
    assert runner.state["agents"]["citizens"]["age"].shape == (256, 1)


def test_r_steps_equals_the_per_step_expression_with_its_gradient():
    """`r_steps` (one broadcast product + row sum + gather, constant Jacobian backward) against autograd through the reference's
    per-step expression `(R2 * one_hot[week]).sum()`: values bit for bit, gradients to float32 rounding, for every R2 shape the
    reference's broadcast accepts (Q2: a (NUM_WEEKS, 1) R2 sums all weeks)"""
    cfg = synthetic_config(256, 3.0, device="cpu")
    runner = Runner(cfg, covid.registry)
    runner.init()
    tr = runner.initializer.transition_function["0"]["new_transmission"]
    gen = torch.Generator().manual_seed(3)
    for shape in ((3, 1), (3,), (1, 3), (1,), ()):
        for n_steps in (21, 10, 5):
            base = torch.rand(shape, generator=gen) * 9 + 0.1
            a, b = base.clone().requires_grad_(True), base.clone().requires_grad_(True)
            weight = torch.rand(n_steps, generator=gen)
            tr.calibrate_R2 = a
            per_step = torch.stack([tr.effective_r(t) for t in range(n_steps)])
            (per_step * weight).sum().backward()
            tr.calibrate_R2 = b
            fused = tr.r_steps(n_steps)
            (fused * weight).sum().backward()
            assert torch.equal(per_step, fused)
            assert a.grad.shape == b.grad.shape and torch.allclose(a.grad, b.grad, rtol=1e-6, atol=1e-6)
    tr.calibrate_R2 = torch.rand(3, 1)                           # no gradient requested: no graph
    assert not tr.r_steps(21).requires_grad


def test_models_sharing_helper_names_do_not_break_each_other(tmp_path):
    """The registry tables are class-level (registry.py:8-14) and both models register a `read_from_file`; whichever
    model builds a runner must get its own helpers, whatever the import order."""
    from agenttorch_b200.core import envs
    from agenttorch_b200.core.registry import Registry
    from agenttorch_b200.models import covid, predator_prey
    from agenttorch_b200.models.covid.substeps import utils
    from agenttorch_b200.populations import write_population_folder

    predator_prey.get_registry()                                   # last registration wins ...
    assert Registry.helpers["initialization"]["read_from_file"] is not utils.read_from_file
    rng = np.random.default_rng(0)
    folder = write_population_folder(str(tmp_path / "pop"), rng.integers(0, 6, 64), (rng.random(64) < 0.1).astype(np.float64),
                                     np.stack([rng.integers(0, 64, 200), rng.integers(0, 64, 200)], 1))
    runner = envs.create(model=covid, population=folder)           # ... but covid's executor re-asserts covid's helpers
    assert Registry.helpers["initialization"]["read_from_file"] is utils.read_from_file
    assert runner.state["agents"]["citizens"]["disease_stage"].shape == (64, 1)
    pp_runner = predator_prey.get_runner(predator_prey.synthetic.synthetic_config(8, 8, 5, 9, device="cpu"))
    pp_runner.init()
    assert Registry.helpers["initialization"]["read_from_file"] is not utils.read_from_file


def test_core_exports_match_the_reference_package():
    """agent_torch/core/__init__.py:1-13: Registry, Runner, Controller, Initializer, VectorizedRunner, vectorized,
    DistributedRunner, launch_distributed_simulation, __version__, distributions, soft helpers."""
    from agenttorch_b200 import core
    from agenttorch_b200.core.vectorization import is_vectorized
    from agenttorch_b200.models import covid
    from agenttorch_b200.models.covid.synthetic import synthetic_config

    for name in ("Registry", "Runner", "Controller", "Initializer", "VectorizedRunner", "vectorized", "DistributedRunner",
                 "launch_distributed_simulation", "__version__", "distributions", "compare", "logical_and", "discrete_sample"):
        assert hasattr(core, name), name

    @core.vectorized(in_dims=(0, None))
    def f(x, y):
        return x + y

    assert is_vectorized(f) and f._vmap_in_dims == (0, None) and f(1, 2) == 3 and not is_vectorized(len)

    runner = core.VectorizedRunner(synthetic_config(128, 4.0, device="cpu"), covid.get_registry())
    runner.init()
    assert runner.is_vectorized and isinstance(runner, core.Runner)
    assert runner.vectorized_functions["transition"] == {"0/new_transmission": True, "1/seirm_progression": True}
    assert runner.vectorized_functions["policy"] == {"0/citizens/make_isolation_decision": True}


def test_packed_snapshot_round_trip_on_host_tensors():
    """PackedSnapshot: compact dtypes, one buffer, restore into live tensors (storage kept); a lossy compact dtype is refused"""
    from agenttorch_b200.core.snapshot import PackedSnapshot

    stage = torch.tensor([[0.0], [1.0], [2.0], [3.0]])
    nst = torch.tensor([[500.0], [3.0], [5.0], [41.0]])
    env = torch.linspace(0, 1, 21)
    snap = PackedSnapshot({"agents": {"citizens": {"disease_stage": stage, "next_stage_time": nst}}, "environment": {"daily_infected": env}},
                          compact={"disease_stage": torch.uint8, "next_stage_time": torch.int16}, pin=False)
    assert snap.payload_bytes() == 4 * 1 + 4 * 2 + 21 * 4
    live = {"agents": {"citizens": {"disease_stage": torch.zeros(4, 1), "next_stage_time": torch.zeros(4, 1)}},
            "environment": {"daily_infected": torch.zeros(21)}}
    keep = live["agents"]["citizens"]["disease_stage"]
    for _ in range(3):                       # alternating staging buffers
        keep.fill_(7.0)
        snap.prefetch("cpu")                 # a no-op off the GPU
        snap.restore_into(live)
        assert live["agents"]["citizens"]["disease_stage"] is keep and torch.equal(keep, stage)
        assert torch.equal(live["agents"]["citizens"]["next_stage_time"], nst)
        assert torch.equal(live["environment"]["daily_infected"], env)
    with pytest.raises(ValueError):
        PackedSnapshot({"a": {"stage": torch.tensor([0.5, 1.0])}}, compact={"stage": torch.uint8}, pin=False)
